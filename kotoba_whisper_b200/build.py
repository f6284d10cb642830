"""Builds libkwb200.so in-tree with nvcc for sm_100a (cross-compiles on a CPU-only box)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libkwb200.so")
SOURCES = ["api.cu", "logmel.cu", "elementwise.cu", "gemm_simt.cu", "gemm_tc.cu", "attention_tc.cu", "attention.cu", "sampling.cu", "decode_fused.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr", "-Xptxas", "-v"] + os.environ.get("KW_NVCC_EXTRA", "").split()


def _newer(src: str, dst: str) -> bool:
    if not os.path.exists(dst):
        return True
    deps = [src, os.path.join(HERE, "..", "include", "kwb200.h")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    return any(os.path.getmtime(d) > os.path.getmtime(dst) for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, variant: str = "", extra: list[str] | None = None) -> str:
    """variant: instrumented side build (objects in build_<variant>/, library libkwb200_<variant>.so, loaded with
    KW_LIB_VARIANT=<variant>); extra: additional nvcc flags (e.g. -DKW_FD_TIMING=1)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build" + ("_" + variant if variant else ""))
    lib_path = os.path.join(HERE, "libkwb200" + ("_" + variant if variant else "") + ".so")
    flags = NVCC_FLAGS + (extra or [])
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(objdir, s.replace(".cu", ".o"))
        if force or _newer(src, obj):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        r = subprocess.run([nvcc, *flags, "-c", src, "-o", obj], capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        with open(obj + ".log", "w") as f:
            f.write(r.stdout + r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(jobs) or 1)) as ex:
        list(ex.map(compile_one, jobs))
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in SOURCES]
    if jobs or not os.path.exists(lib_path):
        r = subprocess.run([nvcc, "-shared", "-o", lib_path, *objs, "-lcudart_static", "-ldl", "-lrt", "-lpthread"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return lib_path


if __name__ == "__main__":
    variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else ""
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=variant,
                extra=[a for a in sys.argv[1:] if a.startswith("-D")]))
