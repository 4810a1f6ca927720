"""Data-parallel step over NCCL (2 ranks) == single-GPU step on the concatenated batch.
Needs two GPUs; skipped otherwise (the single-GPU identity test in test_gpu_model.py and the gloo
test in test_cpu_abi_and_host.py cover the arithmetic and the host logic)."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["REPO"])
from nlpscratch_b200 import Transformer
from nlpscratch_b200.dp import DataParallel, shard_rows
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
cfg = (211, 64, 2, 128, 2, 48)
rng = np.random.default_rng(0)
x = rng.integers(1, 211, (8, 48)).astype(np.int32); x[3, 30:] = 0
y = rng.integers(0, 211, (8, 48)).astype(np.int32)
np.random.seed(11 + rank)                      # ranks start from DIFFERENT weights: broadcast must fix it
model = Transformer(*cfg, dropout_p=0.0, precision="fp32", device=local)
dp = DataParallel(model, min_bucket_floats=1, backend=os.environ.get("DP_BACKEND", "native"))  # one all-reduce per bucket
dp.broadcast_parameters(0)
b, e = shard_rows(8, rank, world)
from nlpscratch_b200.array import DeviceArray
xd, yd = DeviceArray.from_host(x[b:e]), DeviceArray.from_host(y[b:e])
STEPS = 5                                      # device-resident inputs: eager, captured, then replayed as a CUDA graph
for step in range(STEPS):
    loss = dp.train_step(xd, yd, 1e-3, 0.5)
got = model.get_params()
import hashlib
print("rank%d params sha %s" % (rank, hashlib.sha256(b"".join(np.ascontiguousarray(got[k]).tobytes() for k in sorted(got))).hexdigest()))
losses = torch.tensor([float(loss)], device="cuda"); dist.all_reduce(losses)
if rank == 0:
    np.random.seed(11)
    ref = Transformer(*cfg, dropout_p=0.0, precision="fp32", device=local)
    for step in range(STEPS):
        l = ref.train_step(x, y, 1e-3, 0.5)
    want = ref.get_params()
    worst = max(float(np.linalg.norm(got[k].astype(np.float64) - want[k]) / max(np.linalg.norm(want[k]), 1e-30)) for k in want)
    assert worst < 1e-5, worst
    assert abs(float(losses.item()) / world - float(l)) < 1e-4, (losses.item() / world, float(l))
    print("dp ok", worst, dp.launches, dp.backend)
# the unfused API under data parallelism: forward / loss / backward, then the library's whole-buffer exchange
if dp.backend == "native":
    logits, cache = model.forward(xd)
    l2, dl = model.calculate_loss_dlogits(logits, yd, grad_scale=1.0 / world)
    grads = model.backward(cache, dl)
    dp.allreduce_gradients()
    g = model.flat("grad").get()
    digest = hashlib.sha256(g.tobytes()).hexdigest()
    both = [None] * world
    dist.all_gather_object(both, digest)
    assert len(set(both)) == 1, both
dist.barrier(); dp.close(); dist.destroy_process_group()
'''


def test_two_rank_nccl_step_matches_single_gpu(tmp_path):
    from nlpscratch_b200 import _lib
    if _lib.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = os.path.join(tmp_path, "worker.py")
    with open(script, "w") as fh:
        fh.write(WORKER)
    env = dict(os.environ, REPO=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29541", script]
    shas = {}
    # native: NCCL owned by libnlps_b200.so, exchange + clip + Adam inside the one (captured) step call; graph on / off.
    # torch: the round-1 path (dist.all_reduce per bucket from Python; the deferred half-step replayed as a CUDA graph whose
    # bucket events are external event-record nodes, or eager).  All four must give the same bits on both ranks.
    for backend, var, flag in (("native", "NLPS_GRAPH", "1"), ("native", "NLPS_GRAPH", "0"),
                               ("torch", "NLPS_GRAPH_DP", "1"), ("torch", "NLPS_GRAPH_DP", "0")):
        run_env = dict(env, DP_BACKEND=backend)
        run_env[var] = flag
        res = subprocess.run(cmd, env=run_env, capture_output=True, text=True, timeout=280)
        assert res.returncode == 0, (backend, var, flag, res.stdout[-3000:] + res.stderr[-3000:])
        assert "dp ok" in res.stdout
        key = (backend, flag)
        shas[key] = re.findall(r"params sha ([0-9a-f]{64})", res.stdout)     # the ranks' lines may interleave
        assert len(shas[key]) == 2 and shas[key][0] == shas[key][1], shas
    assert len({tuple(v) for v in shas.values()}) == 1, shas
