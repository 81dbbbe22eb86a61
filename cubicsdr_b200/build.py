"""Build cubicsdr_b200/libsdr_b200.so with nvcc for sm_100a (in-tree, so that it travels with the repo snapshot)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "csrc", "engine.cu"), os.path.join(HERE, "csrc", "design.cpp")]
HDR = [os.path.join(HERE, "csrc", f) for f in ("backend.h", "design.h", "dsp_core.h", "dsp_tail.h", "dsp_spec.h", "dsp_pfb.h", "group.h")] + [
    os.path.join(ROOT, "include", "sdr_b200.h")]
OUT = os.path.join(HERE, "libsdr_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
              "-I", os.path.join(ROOT, "include")]


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile the CUDA library.  Raises if nvcc is missing and no prebuilt library is present."""
    if not force and not _stale(OUT, SRC + HDR):
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        if os.path.exists(OUT):
            return OUT
        raise RuntimeError("nvcc not found and %s is missing" % OUT)
    cmd = [nvcc] + NVCC_FLAGS + SRC + ["-o", OUT]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout[-6000:])
    if verbose:
        print(r.stdout)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
