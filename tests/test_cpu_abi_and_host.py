"""CPU-side checks: the C-ABI library builds/loads and exports every declared symbol, the product
refuses to run without a GPU (no fallback), and the data-parallel host logic (sharding, bucket
plan, the sum-of-scaled-shards identity) holds under a world_size-2 gloo group."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from nlpscratch_b200 import _lib
    names = _lib.declared_symbols()
    assert len(names) >= 60 and "nlps_train_step" in names and "nlps_gemm" in names
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), n
    assert lib.nlps_abi_version() == 1
    out = subprocess.run(["nm", "-D", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert set(names) <= exported


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    from nlpscratch_b200 import _lib
    _lib.load()
    res = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True)
    if res.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    sass = res.stdout
    assert "UTCHMMA" in sass or "UTCMMA" in sass      # tcgen05.mma
    assert "UTMALDG" in sass                          # TMA tensor loads
    assert "LDTM" in sass                             # tcgen05.ld


def test_no_gpu_means_loud_failure_not_fallback():
    from nlpscratch_b200 import Transformer, _lib
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Transformer(10, 8, 2, 16, 1, 8)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "nlpscratch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_shard_rows_and_weights():
    from nlpscratch_b200.dp import bucket_plan, shard_rows, shard_weight
    assert [shard_rows(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [shard_rows(16, r, 2) for r in range(2)] == [(0, 8), (8, 16)]
    assert abs(sum(shard_weight(10, r, 4) for r in range(4)) - 1.0) < 1e-12
    assert shard_weight(16, 0, 2) == 0.5
    assert bucket_plan([5, 100, 3, 100, 2], 50) == [(0, 2), (2, 4), (4, 5)]
    assert bucket_plan([10, 10], 1000) == [(0, 2)]


WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["REPO"])
from oracle.transformer_oracle import OracleTransformer, global_norm_clip
from nlpscratch_b200.dp import shard_rows, shard_weight, allreduce_sum_host
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
cfg = (40, 16, 2, 32, 2, 12)
np.random.seed(0)
model = OracleTransformer(*cfg, dropout_p=0.0)
rng = np.random.default_rng(0)
x = rng.integers(1, 40, (5, 12)).astype(np.int32)        # 5 rows over 2 ranks: ragged shards 3 + 2
y = rng.integers(0, 40, (5, 12)).astype(np.int32)
b, e = shard_rows(5, rank, world)
logits, cache = model.forward(x[b:e])
loss, dlogits = model.calculate_loss_dlogits(logits, y[b:e])
w = shard_weight(5, rank, world, 12)
grads = model.backward(cache, dlogits * w)
arrays = [grads[k] for k in sorted(grads)]
allreduce_sum_host(arrays)
# single-process truth on the concatenated batch
np.random.seed(0)
ref = OracleTransformer(*cfg, dropout_p=0.0)
lg, c = ref.forward(x)
l, dl = ref.calculate_loss_dlogits(lg, y)
g = ref.backward(c, dl)
worst = max(float(np.linalg.norm(grads[k].astype(np.float64) - g[k]) / max(np.linalg.norm(g[k]), 1e-30)) for k in g)
assert worst < 1e-5, worst
global_norm_clip([grads[k] for k in grads], 1.0); model.step(grads, 1e-3)
global_norm_clip([g[k] for k in g], 1.0); ref.step(g, 1e-3)
for k in g:
    assert np.allclose(model.params[k], ref.params[k], rtol=0, atol=2e-6), k
# save_checkpoint_rank0 (SURVEY 8f #4): rank 0 alone writes, every rank returns only after the file is complete
from nlpscratch_b200.dp import save_checkpoint_rank0
import pickle, types
class _Model:
    def save_checkpoint(self, filename):
        import time
        time.sleep(0.3)                                    # a slow writer: the barrier must cover it
        with open(filename, "wb") as fh:
            pickle.dump({"params": {k: v for k, v in model.params.items()}, "adam_t": model.adam_t}, fh)
ck = os.path.join(os.environ["CKDIR"], "ck.pkl")
save_checkpoint_rank0(types.SimpleNamespace(rank=rank, model=_Model(), _dist=dist, group=None), ck)
with open(ck, "rb") as fh:                                 # readable on EVERY rank right after the call
    state = pickle.load(fh)
assert state["adam_t"] == 1 and all(np.array_equal(state["params"][k], model.params[k]) for k in model.params)
assert os.path.getsize(ck) > 0 and len([f for f in os.listdir(os.environ["CKDIR"]) if f.endswith(".pkl")]) == 1
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok", worst)
'''


def test_two_rank_gloo_data_parallel_identity(tmp_path):
    script = os.path.join(tmp_path, "worker.py")
    with open(script, "w") as fh:
        fh.write(WORKER)
    env = dict(os.environ, REPO=ROOT, OMP_NUM_THREADS="2", OPENBLAS_NUM_THREADS="2", CKDIR=str(tmp_path))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", script]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=280)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("ok") == 2
