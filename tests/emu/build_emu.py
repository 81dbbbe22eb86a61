"""Build tests/emu/libsdr_b200_emu.so: the engine sources compiled by g++ with -DSDR_EMU.

TEST SCAFFOLDING ONLY.  With SDR_EMU every kernel's phases are replayed serially on the CPU (see
cubicsdr_b200/csrc/backend.h) so that the no-GPU test tier can check index arithmetic, state carry and host logic
against the oracle.  The cubicsdr_b200 package never loads this library; it is not a fallback and is never timed.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "cubicsdr_b200", "csrc")
OUT = os.path.join(HERE, "libsdr_b200_emu.so")


def build(force=False):
    srcs = [os.path.join(CSRC, "engine.cu"), os.path.join(CSRC, "design.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")] + [os.path.join(ROOT, "include", "sdr_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    cmd = ["g++", "-O2", "-ffp-contract=off", "-std=c++14", "-DSDR_EMU", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
           "-x", "c++", srcs[0], srcs[1], "-o", OUT]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
