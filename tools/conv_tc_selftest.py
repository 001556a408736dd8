"""Manual GPU harness: tcgen05 implicit-GEMM conv vs the fp32 SIMT implicit GEMM (lip_selftest_conv_tc)."""
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
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 0
only_big = len(sys.argv) > 2 and sys.argv[2] == "big"
# (imgs, H, W, cin, cout, ksz, stride, batch)
shapes = [(4, 32, 32, 32, 32, 3, 1, 2), (3, 16, 16, 64, 64, 3, 1, 2), (5, 8, 8, 128, 128, 3, 1, 3), (4, 16, 16, 32, 64, 1, 1, 2),
          (2, 32, 32, 64, 32, 3, 1, 1), (7, 8, 8, 32, 32, 3, 1, 2), (3, 32, 32, 32, 64, 3, 2, 2), (5, 16, 16, 64, 128, 3, 2, 2),
          (3, 32, 32, 32, 64, 1, 2, 2), (5, 16, 16, 64, 128, 1, 2, 3)]
if iters > 0:
    shapes += [(100, 32, 32, 32, 32, 3, 1, 32), (100, 16, 16, 64, 64, 3, 1, 32), (100, 8, 8, 128, 128, 3, 1, 32),
               (100, 32, 32, 32, 64, 3, 2, 32), (100, 16, 16, 64, 128, 3, 2, 32)]
if only_big:
    shapes = [sh for sh in shapes if sh[0] == 100]
ok = True
for role in (0, 1, 2, 3):
    for (n, H, W, ci, co, k, sd, b) in shapes:
        err, t1, t2 = C.c_float(-1), C.c_float(0), C.c_float(0)
        rc = L.lip_selftest_conv_tc(role, n, H, W, ci, co, k, sd, b, iters, C.byref(err), C.byref(t1), C.byref(t2), None)
        msg = "" if rc == 0 else L.lip_last_error().decode()
        flop = 2.0 * n * (H // sd) * (W // sd) * k * k * ci * co * b * (2 if role == 0 else 1)
        tf = f" tc {t1.value:.3f} ms ({flop / t1.value / 1e9:.1f} TFLOP/s) simt {t2.value:.3f} ms" if iters > 0 and rc == 0 else ""
        print(f"role {role} imgs={n} {H}x{W} cin={ci} cout={co} k={k} stride={sd} batch={b}: rc={rc} rel_err={err.value:.3e}{tf} {msg}", flush=True)
        if rc != 0 or not (err.value < 5e-6):
            ok = False
        if rc == -2:
            sys.exit(2)
print("ALL OK" if ok else "FAILURES")
