"""Manual GPU harness: tcgen05 3xTF32 GEMM vs the exact SIMT GEMM (lip_selftest_tc_gemm)."""
import ctypes as C
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lip_b200
from lip_b200 import _cabi

L = _cabi.lib()
torch.cuda.init()
torch.zeros(1, device="cuda")
shapes = [(128, 128, 32, 1), (128, 128, 64, 1), (128, 128, 256, 2), (256, 256, 128, 3), (512, 1024, 784, 2),
          (784, 1024, 512, 2), (100, 96, 72, 2), (512, 256, 1024, 4)]
ok = True
for variant in (0, 1, 2):
    for (M, N, K, b) in shapes:
        err = C.c_float(-1)
        rc = L.lip_selftest_tc_gemm(variant, M, N, K, b, C.byref(err), None)
        msg = "" if rc == 0 else L.lip_last_error().decode()
        print(f"variant {variant} M={M} N={N} K={K} batch={b}: rc={rc} rel_err={err.value:.3e} {msg}", flush=True)
        if rc != 0 or not (err.value < 5e-6):
            ok = False
        if rc == -2:
            sys.exit(2)
print("ALL OK" if ok else "FAILURES")
