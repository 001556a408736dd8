"""Builds liblip_b200.so (sm_100a only) in-tree with nvcc.  Used by __graft_entry__.build() and the Makefile."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblip_b200.so")
SOURCES = ["lip_api.cu", "lip_gemm_simt.cu", "lip_gemm_tc.cu", "lip_conv_tc.cu", "lip_model.cu", "lip_cnn.cu", "lip_cnn_fused.cu", "lip_resnet.cu", "lip_vecops.cu", "lip_krylov.cu", "lip_comm.cu", "lip_tridiag.cu", "lip_zgrad.cu", "lip_eval.cu"]
NVCC_FLAGS = [
    "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
    "-std=c++17",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "lip_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles the library if it is missing or older than its sources.  Safe under torchrun: ranks serialise on a file
    lock, the first one builds into a temporary file and renames it into place, the others find it up to date."""
    import fcntl
    if not force and not needs_build():
        return LIB
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                return LIB
            tmp = LIB + f".tmp{os.getpid()}"
            # one nvcc -c per source, in parallel (the two GEMM files dominate), then one link
            from concurrent.futures import ThreadPoolExecutor
            objdir = os.path.join(HERE, "build")
            os.makedirs(objdir, exist_ok=True)
            nvcc = _nvcc()
            cflags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])

            def compile_one(src):
                obj = os.path.join(objdir, os.path.splitext(src)[0] + f".{os.getpid()}.o")
                r = subprocess.run([nvcc] + cflags + ["-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
                return src, obj, r

            with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
                results = list(pool.map(compile_one, SOURCES))
            log = "".join(r.stderr for _, _, r in results)
            failed = [(src, r) for src, _, r in results if r.returncode != 0]
            if failed:
                for _, obj, _ in results:
                    if os.path.exists(obj):
                        os.remove(obj)
                raise RuntimeError("nvcc failed:\n" + "".join(r.stdout + r.stderr for _, r in failed))
            objs = [obj for _, obj, _ in results]
            res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs + ["-ldl"],
                                 capture_output=True, text=True)
            for obj in objs:
                os.remove(obj)
            res.stderr = log + res.stderr
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
            os.replace(tmp, LIB)
            if verbose:
                sys.stderr.write(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
