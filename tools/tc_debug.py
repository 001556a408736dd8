import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["LIP_TC_DEBUG"] = "1"
import torch, lip_b200
from lip_b200 import _cabi
L = _cabi.lib(); torch.zeros(1, device="cuda")
for variant in (0, 1, 2):
    for (M, N, K, b) in [(128, 128, 32, 1), (128, 128, 8, 1)]:
        err = C.c_float(-1)
        rc = L.lip_selftest_tc_gemm(variant, M, N, K, b, C.byref(err), None)
        print(f"variant {variant} M={M} N={N} K={K}: rc={rc} err={err.value:.3e}", flush=True)
