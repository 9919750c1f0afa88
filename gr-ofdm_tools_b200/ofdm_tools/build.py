"""Builds libofdmx.so (sm_100a) in-tree with nvcc.  nvcc cross-compiles without a GPU.

The library is several translation units -- the host logic with the small kernels, and one object per
(big template kernel, fft_len) -- compiled in parallel and linked into one shared object.  Objects are
rebuilt when their source, any header of csrc/ or include/ofdmx.h is newer."""
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(_HERE, "..", ".."))
CSRC = os.path.normpath(os.path.join(_HERE, "..", "csrc"))
LIBDIR = os.path.normpath(os.path.join(_HERE, "..", "lib"))
OBJDIR = os.path.join(LIBDIR, "obj")
# (object name, source, extra defines)
UNITS = [("api", "ofdmx_api.cu", []), ("frame1k", "ofdmx_k_frame1k.cu", []), ("framep", "ofdmx_k_framep.cu", [])] + \
        [("framew_%d" % n, "ofdmx_k_framew.cu", ["-DOFDMX_FW_N=%d" % n]) for n in (1024, 2048, 64, 128, 256, 512)] + \
        [("txw_%d" % n, "ofdmx_k_txw.cu", ["-DOFDMX_TXW_N=%d" % n]) for n in (1024, 64, 128, 256, 512)]
CFLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
          "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _headers():
    hs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    return hs + [os.path.join(ROOT, "include", "ofdmx.h")]


def build(force=False, verbose=False, only=None):
    """only: iterable of object names to (re)compile even if up to date (the others are still checked)."""
    os.makedirs(OBJDIR, exist_ok=True)
    so = os.path.join(LIBDIR, "libofdmx.so")
    nvcc = nvcc_path()
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    hdr_t = max(os.path.getmtime(h) for h in _headers())
    jobs = []
    for name, src, defs in UNITS:
        obj = os.path.join(OBJDIR, name + ".o")
        srcp = os.path.join(CSRC, src)
        stale = force or (only and name in only) or not os.path.exists(obj) or \
            os.path.getmtime(obj) < max(hdr_t, os.path.getmtime(srcp))
        if stale:
            cmd = [nvcc] + CFLAGS + (["-Xptxas", "-v"] if verbose else []) + defs + \
                  ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-c", "-o", obj, srcp]
            jobs.append((name, cmd))

    def run(job):
        name, cmd = job
        r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return name, r.returncode, r.stdout

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            results = list(ex.map(run, jobs))
        for name, rc, out in results:
            if verbose or rc != 0:
                print("---- %s\n%s" % (name, out))
            if rc != 0:
                raise RuntimeError("nvcc failed on %s" % name)
    objs = [os.path.join(OBJDIR, name + ".o") for name, _, _ in UNITS]
    if jobs or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(o) for o in objs):
        subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so] + objs, env=env)
    return so


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
