#!/usr/bin/env python3
"""Compile oracle/liquid_port.c (the plain-C restatement) into oracle/_port/libsdr_port.so.  TEST INFRASTRUCTURE ONLY."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "liquid_port.c")
OUT = os.path.join(HERE, "_port", "libsdr_port.so")


def build(force=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-std=gnu99", "-fPIC", "-shared", SRC, "-o", OUT, "-lm"])
    return OUT


if __name__ == "__main__":
    print(build(force=True))
