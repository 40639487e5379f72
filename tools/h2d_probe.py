"""Pinned H2D bandwidth probe: one big copy vs many scan-sized copies on 1/2/4 streams."""
import torch, time
MB = 1 << 20
tot = 60 * MB
host = torch.empty(tot, dtype=torch.uint8).pin_memory()
dev = torch.empty(tot, dtype=torch.uint8, device="cuda")
def run(nchunks, nstreams, reps=10):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    chunk = tot // nchunks
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        for i in range(nchunks):
            with torch.cuda.stream(streams[i % nstreams]):
                dev[i * chunk:(i + 1) * chunk].copy_(host[i * chunk:(i + 1) * chunk], non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return tot / best / 1e9
for nchunks, ns in ((1, 1), (32, 1), (32, 2), (32, 4), (64, 1), (64, 2), (8, 1)):
    print("chunks %3d streams %d: %.1f GB/s" % (nchunks, ns, run(nchunks, ns)))
