"""CPU, world_size 2 over gloo: the N>1 path (sequence sharding, max-over-ranks timing, pose
gather).  The per-rank work here is the oracle (CPU); on GPUs bench.py runs the CUDA path with
the same plumbing."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from liodom_b200 import sharding


def test_shard_partition_properties():
    for n in (1, 2, 7, 8, 9, 64):
        for world in (1, 2, 3, 4, 8):
            parts = [sharding.shard_sequences(n, world, r) for r in range(world)]
            flat = [i for p in parts for i in p]
            assert flat == list(range(n))                       # disjoint, complete, ordered
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert sharding.seed_of(3) == 1003


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_seq, nframes, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import oracle
    from liodom_b200 import synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = sharding.shard_sequences(n_seq, world, rank)
    op = oracle.make_params(prev_frames=5, omp_threads=1)
    poses = []
    for sid in ids:
        scans, _ = synth.sequence("hdl64_small", sharding.seed_of(sid), nframes)
        p, _, _ = oracle.run_sequence(op, scans)
        poses.append(p)
    fake_ms = 10.0 * (rank + 1)                                   # rank 1 is the slow one
    value, ms = sharding.job_throughput(len(ids) * nframes, fake_ms, world)
    allp = sharding.gather_poses(np.stack(poses), ids, n_seq)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), poses=allp, value=value, ms=ms)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo(tmp_path):
    n_seq, nframes, world = 4, 3, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_seq, nframes, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["poses"], r1["poses"])               # every rank sees the whole job
    assert float(r0["ms"]) == 20.0 and float(r1["ms"]) == 20.0     # max over ranks
    assert np.isclose(float(r0["value"]), world * (n_seq // world) * nframes / 0.020)
    # and it equals the single-process result
    import oracle
    from liodom_b200 import synth
    op = oracle.make_params(prev_frames=5, omp_threads=1)
    for sid in range(n_seq):
        scans, _ = synth.sequence("hdl64_small", sharding.seed_of(sid), nframes)
        p, _, _ = oracle.run_sequence(op, scans)
        assert np.array_equal(r0["poses"][sid], p)
