#!/bin/bash
# tests/emu/asan_run.sh -- TEST SCAFFOLDING: the no-GPU tier under AddressSanitizer.
# Builds the kernel emulator (the engine's host logic + every kernel's phases replayed serially, -DSDR_EMU) with
# -fsanitize=address, puts it where the emu_lib fixture looks, runs the CPU tier, and rebuilds the ordinary library.
# new_delete_type_mismatch=0: the reference's own ModemCW::disposeKit deletes its kit through the base pointer.
set -e
cd "$(dirname "$0")/../.."
ASAN=$(gcc -print-file-name=libasan.so)
g++ -O1 -g -fsanitize=address -fno-omit-frame-pointer -ffp-contract=off -std=c++14 -DSDR_EMU -fPIC -shared -I include \
    -x c++ cubicsdr_b200/csrc/engine.cu cubicsdr_b200/csrc/design.cpp -o tests/emu/libsdr_b200_emu.so
trap 'python -c "from tests.emu import build_emu; build_emu.build(force=True)"' EXIT
LD_PRELOAD=$ASAN ASAN_OPTIONS=detect_leaks=0:halt_on_error=1:new_delete_type_mismatch=0 python -m pytest tests -q -m "not gpu" "$@"
