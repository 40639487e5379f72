"""GPU (>= 2 devices): point-sharded mode — ring-sharded extraction, edge-sharded association and
solve with a 29-double NCCL all-reduce per LM evaluation — against the single-GPU path
(tests/run_sharded_check.py under torchrun).  Skipped on single-GPU boxes."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("sensor,regions,frames", [("hdl64_small", 8, 5), ("hdl64", 16, 3)])
def test_point_sharded_matches_single_gpu(cuda_lib, sensor, regions, frames):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "run_sharded_check.py"), "--sensor", sensor,
           "--frames", str(frames), "--scan-regions", str(regions)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    res = json.loads(line)
    assert res["status"] == "ok" and res["world"] == 2
    assert res["worst_pose_diff_m"] < 1e-9 and res["worst_rot_diff_rad"] < 1e-10


def test_shard_init_argument_checks(cuda_lib):
    import numpy as np
    from liodom_b200 import api
    ctx = api.Context(batch=2, max_points=2048)
    with pytest.raises(api.LiodomError):
        ctx.shard_init(0, 2, bytes(128))          # batch-1 contexts only
    ctx.close()
    ctx = api.Context(batch=1, max_points=2048)
    with pytest.raises(api.LiodomError):
        ctx.shard_init(0, 3, bytes(128))          # 64 scan lines are not a multiple of 3
    ctx.shard_init(0, 1, bytes(128))              # world 1: plain single-GPU context
    ctx.close()
