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
SOURCES = ["conv_tc.cu", "conv_pers_v1.cu", "enc_head.cu", "bn_train.cu", "stem_tc.cu", "misc.cu", "attn.cu", "mlp.cu", "wgrad.cu", "bn_bwd.cu",
           "attn_bwd.cu", "bwd_misc.cu"]
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


def _compile_flags():
    return [f for f in NVCC_FLAGS if f != "-shared"]


def _object_for(name, extra_flags, obj_dir):
    """Object file of one source, named by the hash of that source, every header and the flags: editing one .cu
    recompiles one object (the link takes a second), editing a header recompiles all."""
    h = hashlib.sha256()
    for dep in [name] + HEADERS:
        with open(os.path.join(CSRC, dep), "rb") as f:
            h.update(f.read())
    h.update(" ".join(_compile_flags() + list(extra_flags)).encode())
    return os.path.join(obj_dir, "%s.%s.o" % (os.path.splitext(name)[0], h.hexdigest()[:16]))


def _compile_and_link(extra_flags, out, verbose=False):
    """Compile the sources as parallel nvcc processes (one object each, cached under lib/obj) and link them."""
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    objs, procs = [], []
    for name in SOURCES:
        obj = _object_for(name, extra_flags, obj_dir)
        objs.append(obj)
        if not os.path.exists(obj):
            tmp = obj + ".tmp%d" % os.getpid()
            cmd = [nvcc] + _compile_flags() + list(extra_flags) + ["-c", os.path.join(CSRC, name), "-o", tmp]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            procs.append((name, obj, tmp, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    errors = []
    for name, obj, tmp, pr in procs:
        log, _ = pr.communicate()
        if pr.returncode != 0:
            errors.append("%s:\n%s" % (name, log))
        else:
            os.replace(tmp, obj)
            if verbose and log:
                print(log, file=sys.stderr)
    if errors:
        raise RuntimeError("nvcc failed:\n" + "\n".join(errors))
    # drop objects of older source versions
    keep = {os.path.basename(o) for o in objs}
    for f in os.listdir(obj_dir):
        if f.endswith(".o") and f not in keep:
            os.remove(os.path.join(obj_dir, f))
    cmd = [nvcc] + NVCC_FLAGS + objs + ["-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + "\n" + res.stderr)
    return out


def build(force=False, verbose=False, extra_flags=(), out=None):
    """Compile every CUDA source for sm_100a into lib/libw2c.so (or `out`: a variant build, e.g. the instrumented one
    tools/time_enc_head.py makes with -DW2C_HEAD_TIMING; variants are never stamped). Returns the library path."""
    os.makedirs(LIB_DIR, exist_ok=True)
    if out is not None:
        return _compile_and_link(extra_flags, out, verbose)
    fp = _fingerprint(extra_flags)
    if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == fp:
                return LIB_PATH
    if force:
        shutil.rmtree(os.path.join(LIB_DIR, "obj"), ignore_errors=True)
    _compile_and_link(extra_flags, LIB_PATH, verbose)
    with open(STAMP, "w") as f:
        f.write(fp)
    return LIB_PATH


if __name__ == "__main__":
    flags = [a for a in sys.argv[1:] if a not in ("--force", "-v")]
    print(build(force="--force" in sys.argv, verbose=True, extra_flags=flags))
