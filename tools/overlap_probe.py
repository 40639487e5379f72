"""H2D copy bandwidth while other work runs on the GPU (diagnostic for the e2e pipeline)."""
import sys, time, torch
sys.path.insert(0, '/root/repo')
dev = torch.device("cuda", 0)
N = 60 * 1024 * 1024  # floats = 240 MB
host = [torch.empty(N, dtype=torch.float32).pin_memory() for _ in range(4)]
dst = torch.empty(N, dtype=torch.float32, device=dev)
cs = torch.cuda.Stream()
def copy_bw(label, background=None, reps=8):
    torch.cuda.synchronize()
    stop = False
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if background: background(40)
    with torch.cuda.stream(cs):
        e0.record()
        for r in range(reps): dst.copy_(host[r % 4], non_blocking=True)
        e1.record()
    torch.cuda.synchronize()
    print("%-28s H2D %.1f GB/s" % (label, reps * N * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9))
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16); b = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
x = torch.randn(256 * 1024 * 1024, device=dev); y = torch.empty_like(x)
idx = torch.randint(0, x.numel(), (64 * 1024 * 1024,), device=dev)
def bg_matmul(n):
    for _ in range(n): torch.matmul(a, b)
def bg_stream(n):
    for _ in range(n * 4): y.copy_(x)
def bg_gather(n):
    for _ in range(n): torch.index_select(x, 0, idx)
copy_bw("alone")
copy_bw("with bf16 matmul", bg_matmul)
copy_bw("with device copy (HBM bound)", bg_stream)
copy_bw("with random gather", bg_gather)
