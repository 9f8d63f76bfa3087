"""Build libw2c.so (the sm_100a CUDA kernels + C ABI) in-tree with nvcc.

nvcc cross-compiles without a GPU, so this runs on the CPU-only build box; the resulting .so is git-ignored but
travels to the GPU box with the repo snapshot.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libw2c.so")
STAMP = os.path.join(LIB_DIR, "libw2c.stamp")
SOURCES = ["conv_tc.cu", "conv_pers_v1.cu", "enc_head.cu", "bn_train.cu", "stem_tc.cu", "misc.cu", "attn.cu", "mlp.cu"]
HEADERS = ["ptx.cuh", "common.cuh", "conv_plan.cuh", os.path.join("..", "..", "include", "w2c.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "static",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _fingerprint(extra_flags):
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS + list(extra_flags)).encode())
    return h.hexdigest()


def is_current(extra_flags=()):
    """True when lib/libw2c.so exists and its stamp equals the fingerprint of the present sources and flags."""
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as f:
        return f.read().strip() == _fingerprint(extra_flags)


def build(force=False, verbose=False, extra_flags=(), out=None):
    """Compile every CUDA source for sm_100a into lib/libw2c.so (or `out`: a variant build, e.g. the instrumented one
    tools/time_enc_head.py makes with -DW2C_HEAD_TIMING; variants are never stamped). Returns the library path."""
    os.makedirs(LIB_DIR, exist_ok=True)
    if out is not None:
        cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", out]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + "\n" + res.stderr)
        return out
    fp = _fingerprint(extra_flags)
    if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == fp:
                return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB_PATH]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + "\n" + res.stderr)
    if verbose and (res.stdout or res.stderr):
        print(res.stdout + res.stderr, file=sys.stderr)
    with open(STAMP, "w") as f:
        f.write(fp)
    return LIB_PATH


if __name__ == "__main__":
    flags = [a for a in sys.argv[1:] if a not in ("--force", "-v")]
    print(build(force="--force" in sys.argv, verbose=True, extra_flags=flags))
