import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import test_gpu_kernels as T
B,S,h,d,pad = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5])
T.test_attention_forward_backward(B,S,h,d,pad,"tf32")
print("ok", B,S,h,d,pad)
