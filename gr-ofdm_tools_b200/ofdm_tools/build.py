"""Builds libofdmx.so (sm_100a) in-tree with nvcc.  nvcc cross-compiles without a GPU."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(_HERE, "..", ".."))
CSRC = os.path.normpath(os.path.join(_HERE, "..", "csrc"))
LIBDIR = os.path.normpath(os.path.join(_HERE, "..", "lib"))
SOURCES = ["ofdmx_api.cu"]
HEADERS = ["ofdmx_dev.cuh", "ofdmx_kernels.cuh", "ofdmx_sync.cuh", "ofdmx_frame1024.cuh", "fft32_gen.cuh", "ofdmx_chain.cuh", "ofdmx_sync_tma.cuh", "ofdmx_frame1024w.cuh", "ofdmx_cond.cuh", "ofdmx_sync_warp.cuh", "ofdmx_tx1024w.cuh", "ofdmx_symbol_small.cuh", "fft_small_gen.cuh"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    so = os.path.join(LIBDIR, "libofdmx.so")
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(ROOT, "include", "ofdmx.h")]
    if not force and os.path.exists(so) and all(os.path.getmtime(d) <= os.path.getmtime(so) for d in deps):
        return so
    cmd = [nvcc_path()] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", so] + [os.path.join(CSRC, f) for f in SOURCES]
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    return so


if __name__ == "__main__":
    print(build(force=True, verbose=True))
