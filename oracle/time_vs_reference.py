"""Times the real reference (imported from /root/reference, this container only) and the oracle port on the same
c4 sample (batch 4 x seq 512), so that the port can stand in for the reference as the CPU baseline on the GPU box.
Measured here (8 cores, numpy 2.3.5 / OpenBLAS): reference 343 tokens/s, oracle port 334 tokens/s."""
import sys, time, numpy as np
sys.path.insert(0, "/root/reference"); sys.path.insert(0, "/root/repo")
sys.dont_write_bytecode = True
from Transformer.transformer import Transformer as Ref
from utils.function import clip_grads
from oracle.transformer_oracle import OracleTransformer, full_sequence_step
L,E,h,F,S,B,V = 6,512,8,2048,512,4,8192
rng = np.random.default_rng(1234)
x = rng.integers(1, V, (B, S)).astype(np.int32); y = rng.integers(1, V, (B, S)).astype(np.int32)
np.random.seed(0)
ref = Ref(V,E,h,F,L,S, dropout_p=0.1)
def ref_step():
    logits, cache = ref.forward(x)
    loss, dl = ref.calculate_loss_dlogits(logits, y)
    g = ref.backward(cache, dl)
    clip_grads(list(g.values()), 1.0, lib=np)
    ref.step(g, 1e-4)
ref_step()
t=time.perf_counter(); ref_step(); ref_step(); dt=(time.perf_counter()-t)/2
print("reference", dt, B*S/dt)
np.random.seed(0)
o = OracleTransformer(V,E,h,F,L,S, dropout_p=0.1)
full_sequence_step(o,x,y,lr=1e-4,clip=1.0)
t=time.perf_counter(); full_sequence_step(o,x,y,lr=1e-4,clip=1.0); full_sequence_step(o,x,y,lr=1e-4,clip=1.0); dt=(time.perf_counter()-t)/2
print("oracle", dt, B*S/dt)
