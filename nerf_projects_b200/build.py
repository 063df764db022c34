"""Builds libnerf_b200.so in-tree with nvcc for sm_100a (no torch extension
machinery: the library is a plain C ABI, loaded with ctypes)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# experiments: NERFB_VARIANT=name NERFB_DEFINES="-DX=1" builds libnerf_b200_<name>.so beside the product library
VARIANT = os.environ.get("NERFB_VARIANT", "")
LIB = os.path.join(HERE, f"libnerf_b200_{VARIANT}.so" if VARIANT else "libnerf_b200.so")
OBJ_DIR = os.path.join(HERE, "build", VARIANT) if VARIANT else os.path.join(HERE, "build")

SOURCES = ["api.cu", "rays.cu", "embed.cu", "linear_fp32.cu", "composite.cu", "sample_pdf.cu", "adam.cu", "rng.cu", "metrics.cu",
           "umma_selftest.cu", "mlp_tc.cu", "wgrad_tc.cu"]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"] + os.environ.get("NERFB_DEFINES", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stamp(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(os.path.basename(p).encode()); h.update(f.read())     # content-addressed: independent of the checkout path
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "nerf_b200.h"))
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    # experiments: NERFB_SRC_OVERRIDE_DIR holds alternative versions of some sources (A/B against an older kernel)
    over = os.environ.get("NERFB_SRC_OVERRIDE_DIR")
    if over:
        srcs = [os.path.join(over, os.path.basename(p)) if os.path.exists(os.path.join(over, os.path.basename(p))) else p
                for p in srcs]
    nvcc = _nvcc()
    objs, jobs = [], []
    for src in srcs:
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        stamp_file = obj + ".stamp"
        stamp = _stamp([src] + headers)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
            continue
        jobs.append((src, obj, stamp_file, stamp))

    def compile_one(job):
        src, obj, stamp_file, stamp = job
        # -I csrc: override sources (NERFB_SRC_OVERRIDE_DIR) live elsewhere but include the in-tree headers; the include
        # path is not part of the stamp, so a checkout at another absolute path does not rebuild
        cmd = [nvcc] + NVCC_FLAGS + ["-I", CSRC, "-c", src, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log = res.stdout + res.stderr
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + log)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{log}")
        with open(stamp_file, "w") as f:
            f.write(stamp)
        return log

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for log in ex.map(compile_one, jobs):
                if verbose:
                    print(log)
    if jobs or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
