"""Manual GPU harness: tensor-core GEMM microbenchmark with diagnosis knobs + SM clock sampling."""
import ctypes as C, os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, lip_b200
from lip_b200 import _cabi
L = _cabi.lib(); torch.zeros(1, device="cuda")
clk = []
stop = False
def sample():
    while not stop:
        try:
            o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip().split(",")
            clk.append((time.time(), float(o[0]), float(o[1])))
        except Exception:
            pass
        time.sleep(0.05)
th = threading.Thread(target=sample, daemon=True); th.start()
shapes = {"jvp_l1": (0, 512, 512, 2048, 256), "wgrad_l1": (1, 1024, 512, 512, 256), "dgrad_l1": (2, 512, 1024, 512, 256)}
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for name, (v, M, N, K, b) in shapes.items():
    for two in (0, 1, 2):
        for dbg in (0, 1, 5, 7):
            ms = C.c_float(0)
            t0 = time.time()
            rc = L.lip_bench_tc_gemm(v, M, N, K, b, iters, dbg, two, C.byref(ms), None)
            t1 = time.time()
            if rc != 0:
                print(name, two, dbg, "ERROR", L.lip_last_error().decode()); continue
            c = [x for x in clk if t0 <= x[0] <= t1]
            mhz = sorted(x[1] for x in c)[len(c) // 2] if c else -1
            pw = max((x[2] for x in c), default=-1)
            tf = 2.0 * M * N * K * b / (ms.value * 1e-3) / 1e12
            kb = (K + 31) // 32 * ((M + 127) // 128) * ((N + 127) // 128) * b / 148
            print(f"{name:9s} two_cta={two} dbg={dbg}: {ms.value:8.3f} ms  {tf:7.1f} TFLOP/s(fp32-equiv)  {ms.value*1e3/kb*1000:6.0f} ns/k-block/SM  clk~{mhz:.0f} MHz  {pw:.0f} W", flush=True)
stop = True
