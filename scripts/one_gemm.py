"""One tcgen05 projection at a given shape (for ncu captures): python scripts/one_gemm.py M K N"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshrcnn_b200 import functional as F_, _lib
M, K, N = (int(x) for x in sys.argv[1:4])
a = torch.randn(M, K, device="cuda"); w = torch.randn(K, N, device="cuda"); c = torch.empty(M, N, device="cuda")
img = F_.tc_pack(w, None, N, 1, 0, 0, K, N)
for _ in range(5):
    F_.tc_gemm(_lib.ptr(a), K, M, K, img, N, _lib.ptr(c), N)
torch.cuda.synchronize()
print("rel err vs fp64:", float(((c.double() - a.double() @ w.double()).norm() / (a.double() @ w.double()).norm())))
