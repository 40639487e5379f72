"""Point-sharded mode check (needs >= 2 GPUs; launched by torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/run_sharded_check.py [--sensor hdl64_1m] [--frames 4] [--scan-regions 64]

Every rank runs (a) the ordinary single-GPU path and (b) the point-sharded path (ring-sharded
extraction, edge-sharded association/solve, 29-double NCCL all-reduce per LM evaluation) on the same
scans and checks: identical edges, poses within 1e-9 m / 1e-10 rad of the single-GPU result and
bitwise identical across ranks.  Prints one JSON line on rank 0 with the device time per scan.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from liodom_b200 import api, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sensor", default="hdl64_1m")
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--scan-regions", type=int, default=64)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    scans, _ = synth.sequence(a.sensor, 1000, a.frames)
    kw = dict(prev_frames=15, scan_regions=a.scan_regions, max_points=max(32768, 1 << int(np.ceil(np.log2(max(len(s) for s in scans))))), device=local)
    uid = [api.shard_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    single = api.Context(batch=1, **kw)
    shard = api.Context(batch=1, **kw)
    shard.shard_init(rank, world, uid[0])
    stream = torch.cuda.ExternalStream(shard.stream)
    worst_t = worst_r = 0.0
    ms = []
    nedges = []
    for f, s in enumerate(scans):
        single.scan_batch([s])
        p1, n1 = single.results()
        e1 = single.scan_edges(0)
        dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        shard.scan_batch([s])
        ev1.record(stream)
        p2, n2 = shard.results()
        e2 = shard.scan_edges(0)
        ms.append(ev0.elapsed_time(ev1))
        assert n1[0] == n2[0] and np.array_equal(e1.view(np.uint32), e2.view(np.uint32)), "frame %d: edges differ" % f
        dt = np.abs(p1[0][:3, 3] - p2[0][:3, 3]).max()
        dR = p1[0][:3, :3] @ p2[0][:3, :3].T
        ang = 0.5 * np.linalg.norm([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]])
        worst_t, worst_r = max(worst_t, dt), max(worst_r, ang)
        assert dt < 1e-9 and ang < 1e-10, "frame %d: sharded pose differs (%g m, %g rad)" % (f, dt, ang)
        # all ranks hold bitwise the same pose
        t = torch.from_numpy(p2[0].copy()).cuda()
        ref = t.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(t, ref), "frame %d: ranks disagree" % f
        nedges.append(int(n2[0]))
        d = shard.scan_diag(0)
        ds = single.scan_diag(0)
        if f > 0:
            assert d.n_matches[0] == ds.n_matches[0], (d.n_matches[0], ds.n_matches[0])
    t = torch.tensor([float(np.mean(ms[1:]))], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"check": "point_sharded", "world": world, "sensor": a.sensor, "scan_regions": a.scan_regions,
                          "points_per_scan": int(np.mean([len(s) for s in scans])), "edges_per_scan": int(np.mean(nedges)),
                          "ms_per_scan_sharded": round(float(t.item()), 4), "worst_pose_diff_m": float(worst_t),
                          "worst_rot_diff_rad": float(worst_r), "allreduce_doubles": 29, "status": "ok"}))
    single.close()
    shard.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
