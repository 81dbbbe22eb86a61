#!/usr/bin/env python3
"""Build oracle/_ref/ : the reference's OWN liquid-dsp 1.5.0 binary, made runnable on Linux.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is shipped or measured as the product; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use it.

The arithmetic of CubicSDR's hot path lives in liquid-dsp 1.5.0, which the reference vendors only as
a prebuilt MinGW static archive (/root/reference/external/liquid-dsp/gcc/64/libliquid.a, PE/COFF
x86-64 objects, with symbols).  This script never copies reference sources; it reads that archive
where it lies and writes only into oracle/_ref/ (git-ignored, but shipped to the GPU box):

  1. `ar x` the archive into oracle/_ref/obj/
  2. objcopy every COFF object to ELF64, dropping the SEH sections and renaming every libc/libm
     import S to ms_S (the imports follow the Microsoft x64 calling convention)
  3. repair RELA addends: objcopy keeps COFF's in-place addends and writes r_addend = 0, so for
     PC32/PLT32 set r_addend = site - 4, for absolute relocs r_addend = site, and zero the site
  4. compile ms_abi shims that forward ms_S to glibc (oracle/ms_shims.c)
  5. compile the CubicSDR glue restatement (oracle/ref_chain.cpp) and link everything into
     oracle/_ref/libsdr_ref.so

  6. compile the reference's own, unmodified modem sources (src/modules/modem/Modem.cpp, ModemAnalog.cpp,
     analog/Modem*.cpp) from where they lie, their liquid calls renamed to System V forwarders
     (oracle/modem_abi.h, oracle/ref_modem_abi.c), with the C entry points of oracle/ref_modems.cpp, into
     oracle/_ref/libmodems_ref.so (pins ref_chain.cpp's modem restatement: tests/test_oracle_modems.py)
  7. likewise src/process/FFTDataDistributor.cpp, SpectrumVisualProcessor.cpp and ScopeVisualProcessor.cpp, unmodified,
     with the two headers that need wx (the application, the OpenGL panel) shadowed by oracle/shadow/, and the C entry
     points of oracle/ref_visuals.cpp, into oracle/_ref/libvisuals_ref.so (tests/test_oracle_visuals.py)
  8. the whole demodulator: src/demod/DemodulatorInstance.cpp with its DemodulatorPreThread / DemodulatorWorkerThread /
     DemodulatorThread, the audio thread and sinks, and the modems -- and the ingress in front of them,
     src/sdr/SDRPostThread.cpp (SoapySDR and AppConfig headers shadowed: only declarations need them) -- unmodified,
     application header shadowed, with the C entry points of oracle/ref_demod.cpp, into oracle/_ref/libdemod_ref.so
     (tests/test_oracle_demod.py); the same library exposes the audio mixer (the RtAudio callback of
     src/audio/AudioThread.cpp, on the reference's RtAudio built with its dummy API: no device) and the WAV writer
     (src/audio/AudioFile.cpp + AudioFileWAV.cpp) through oracle/ref_audio.cpp

If /root/reference is absent (GPU box) and oracle/_ref/libsdr_ref.so already exists, nothing is done.
"""
import os
import struct
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_AR = "/root/reference/external/liquid-dsp/gcc/64/libliquid.a"
OUT = os.path.join(HERE, "_ref")
OBJ = os.path.join(OUT, "obj")
LIB = os.path.join(OUT, "libsdr_ref.so")
MODEMS_LIB = os.path.join(OUT, "libmodems_ref.so")
VISUALS_LIB = os.path.join(OUT, "libvisuals_ref.so")
REF_SRC = "/root/reference/src"
REF_VISUAL_SRCS = ["FFTDataDistributor.cpp", "SpectrumVisualProcessor.cpp", "ScopeVisualProcessor.cpp"]
DEMOD_LIB = os.path.join(OUT, "libdemod_ref.so")
REF_DEMOD_SRCS = ["demod/DemodulatorInstance.cpp", "demod/DemodulatorPreThread.cpp", "demod/DemodulatorWorkerThread.cpp",
                  "demod/DemodulatorThread.cpp", "audio/AudioSinkThread.cpp", "audio/AudioSinkFileThread.cpp",
                  "audio/AudioFile.cpp", "audio/AudioFileWAV.cpp", "IOThread.cpp", "util/Timer.cpp", "sdr/SDRPostThread.cpp"]
REF_MODEM_SRCS = ["Modem.cpp", "ModemAnalog.cpp"] + ["analog/Modem%s.cpp" % m for m in
                                                     ("FM", "NBFM", "FMStereo", "AM", "CW", "LSB", "USB", "DSB", "IQ")]

# archive-undefined minus archive-defined (functions); these become ms_<name> shims
IMPORT_FUNCS = """___chkstk_ms __mingw_vfprintf __mingw_vfscanf __mingw_vsprintf atan2f calloc ccosf cos cosf
coshf csinf csqrtf div erfcf erff exp expf fclose feof fopen free fwrite hypot lgamma lgammaf log log10 log10f
logf malloc memcpy memmove memset pow qsort rand realloc sin sincosf sinf sinhf strcmp strlen tanf tanhf""".split()
IMPORT_DATA = ["__imp___acrt_iob_func", "__imp__assert"]

R_X86_64_64, R_X86_64_PC32, R_X86_64_PLT32, R_X86_64_32, R_X86_64_32S = 1, 2, 4, 10, 11


def run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        sys.stderr.write("FAILED: %s ...\n%s\n" % (" ".join(cmd)[:300], r.stdout[-4000:]))
        raise SystemExit(1)
    return r.stdout


def fix_rela_addends(path):
    """objcopy COFF->ELF leaves the addend in the relocated field; move it into r_addend."""
    with open(path, "rb") as f:
        b = bytearray(f.read())
    assert b[:4] == b"\x7fELF" and b[4] == 2 and b[5] == 1, path
    e_shoff = struct.unpack_from("<Q", b, 0x28)[0]
    e_shentsize, e_shnum = struct.unpack_from("<HH", b, 0x3A)
    secs = []
    for i in range(e_shnum):
        (name, typ, flags, addr, off, size, link, info, align, entsize) = struct.unpack_from(
            "<IIQQQQIIQQ", b, e_shoff + i * e_shentsize)
        secs.append(dict(type=typ, off=off, size=size, info=info, entsize=entsize))
    nfix = 0
    for s in secs:
        if s["type"] != 4:  # SHT_RELA
            continue
        tgt = secs[s["info"]]
        for k in range(s["size"] // 24):
            ro = s["off"] + 24 * k
            r_offset, r_info, r_addend = struct.unpack_from("<QQq", b, ro)
            rtype = r_info & 0xFFFFFFFF
            site = tgt["off"] + r_offset
            if rtype in (R_X86_64_PC32, R_X86_64_PLT32):
                v = struct.unpack_from("<i", b, site)[0]
                struct.pack_into("<q", b, ro + 16, r_addend + v - 4)
                struct.pack_into("<i", b, site, 0)
            elif rtype == R_X86_64_64:
                v = struct.unpack_from("<q", b, site)[0]
                struct.pack_into("<q", b, ro + 16, r_addend + v)
                struct.pack_into("<q", b, site, 0)
            elif rtype in (R_X86_64_32, R_X86_64_32S):
                v = struct.unpack_from("<i", b, site)[0]
                struct.pack_into("<q", b, ro + 16, r_addend + v)
                struct.pack_into("<i", b, site, 0)
            else:
                raise SystemExit("unexpected relocation type %d in %s" % (rtype, path))
            nfix += 1
    with open(path, "wb") as f:
        f.write(b)
    return nfix


def transplant():
    os.makedirs(OBJ, exist_ok=True)
    run(["ar", "x", REF_AR], cwd=OBJ)
    redef = []
    for s in IMPORT_FUNCS:
        redef += ["--redefine-sym", "%s=ms_%s" % (s, s)]
    elf = []
    for fn in sorted(os.listdir(OBJ)):
        if not fn.endswith(".o") or fn.endswith(".elf.o"):
            continue
        src = os.path.join(OBJ, fn)
        dst = os.path.join(OBJ, fn[:-2] + ".elf.o")
        run(["objcopy", "-O", "elf64-x86-64", "--wildcard", "-R", ".pdata*", "-R", ".xdata*", "-R", ".rdata$zzz"]
            + redef + [src, dst])
        fix_rela_addends(dst)
        os.remove(src)
        elf.append(dst)
    for fn in os.listdir(OBJ):  # stray non-object members (*.rdata)
        p = os.path.join(OBJ, fn)
        if not fn.endswith(".elf.o"):
            os.remove(p)
    return elf


def build(force=False):
    have_ref = os.path.exists(REF_AR)
    if not have_ref:
        if os.path.exists(LIB):
            return LIB
        raise SystemExit("oracle/_ref missing and /root/reference absent: cannot build the reference oracle")
    srcs = [os.path.join(HERE, "ms_shims.c"), os.path.join(HERE, "ref_chain.cpp"), os.path.join(HERE, "liquid_ms.h"),
            os.path.abspath(__file__)]
    if (not force) and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    objs = [os.path.join(OBJ, f) for f in sorted(os.listdir(OBJ))] if os.path.isdir(OBJ) else []
    objs = [o for o in objs if o.endswith(".elf.o")]
    if len(objs) < 160:
        objs = transplant()
    shim_o = os.path.join(OUT, "ms_shims.o")
    chain_o = os.path.join(OUT, "ref_chain.o")
    run(["gcc", "-O2", "-fPIC", "-c", os.path.join(HERE, "ms_shims.c"), "-o", shim_o])
    run(["g++", "-O2", "-std=c++14", "-fPIC", "-pthread", "-I", HERE, "-c", os.path.join(HERE, "ref_chain.cpp"), "-o", chain_o])
    run(["g++", "-shared", "-o", LIB, chain_o, shim_o] + objs +
        ["-lm", "-pthread", "-Wl,-Bsymbolic", "-Wl,-z,notext", "-Wl,--allow-multiple-definition", "-Wl,-z,execstack"])
    return LIB


def build_modems(force=False):
    """oracle/_ref/libmodems_ref.so: the reference's unmodified modem classes over its own liquid binary."""
    if not os.path.exists(REF_SRC):
        if os.path.exists(MODEMS_LIB):
            return MODEMS_LIB
        raise SystemExit("oracle/_ref/libmodems_ref.so missing and /root/reference absent")
    lib = build(force)
    mine = [os.path.join(HERE, f) for f in ("modem_abi.h", "ref_modem_abi.c", "ref_modems.cpp")] + [os.path.abspath(__file__)]
    if (not force) and os.path.exists(MODEMS_LIB) and all(os.path.getmtime(MODEMS_LIB) >= os.path.getmtime(s) for s in mine + [lib]):
        return MODEMS_LIB
    mdir = os.path.join(OUT, "modem_obj")
    os.makedirs(mdir, exist_ok=True)
    mod = os.path.join(REF_SRC, "modules", "modem")
    inc = []
    for d in (REF_SRC, os.path.join(REF_SRC, "util"), os.path.join(REF_SRC, "audio"), os.path.join(REF_SRC, "demod"), mod,
              os.path.join(mod, "analog"), "/root/reference/external/liquid-dsp/include", "/root/reference/external/rtaudio"):
        inc += ["-I", d]
    cxx = ["g++", "-std=c++17", "-O2", "-fPIC", "-pthread", "-w", "-include", os.path.join(HERE, "modem_abi.h")] + inc
    objs = []
    for src in REF_MODEM_SRCS:                                   # compiled where they lie, objects only into oracle/_ref/
        o = os.path.join(mdir, os.path.basename(src)[:-4] + ".o")
        run(cxx + ["-c", os.path.join(mod, src), "-o", o])
        objs.append(o)
    drv = os.path.join(mdir, "ref_modems.o")
    run(cxx + ["-c", os.path.join(HERE, "ref_modems.cpp"), "-o", drv])
    abi = os.path.join(mdir, "ref_modem_abi.o")
    run(["gcc", "-O2", "-fPIC", "-c", os.path.join(HERE, "ref_modem_abi.c"), "-o", abi])
    run(["g++", "-shared", "-o", MODEMS_LIB, drv, abi] + objs +
        ["-L", OUT, "-l:libsdr_ref.so", "-Wl,-rpath,$ORIGIN", "-lm", "-pthread", "-Wl,--no-undefined"])
    return MODEMS_LIB


def build_visuals(force=False):
    """oracle/_ref/libvisuals_ref.so: the reference's unmodified FFTDataDistributor / SpectrumVisualProcessor /
    ScopeVisualProcessor over its own liquid binary."""
    if not os.path.exists(REF_SRC):
        if os.path.exists(VISUALS_LIB):
            return VISUALS_LIB
        raise SystemExit("oracle/_ref/libvisuals_ref.so missing and /root/reference absent")
    lib = build(force)
    shadow = os.path.join(HERE, "shadow")
    mine = [os.path.join(HERE, f) for f in ("modem_abi.h", "ref_modem_abi.c", "ref_visuals.cpp")] + \
           [os.path.join(shadow, f) for f in ("CubicSDR.h", "ScopePanel.h")] + [os.path.abspath(__file__)]
    if (not force) and os.path.exists(VISUALS_LIB) and all(os.path.getmtime(VISUALS_LIB) >= os.path.getmtime(s) for s in mine + [lib]):
        return VISUALS_LIB
    vdir = os.path.join(OUT, "visual_obj")
    os.makedirs(vdir, exist_ok=True)
    proc = os.path.join(REF_SRC, "process")
    inc = []
    for d in (shadow, REF_SRC, os.path.join(REF_SRC, "util"), os.path.join(REF_SRC, "audio"), os.path.join(REF_SRC, "demod"), proc,
              os.path.join(REF_SRC, "modules", "modem"), "/root/reference/external/liquid-dsp/include", "/root/reference/external/rtaudio"):
        inc += ["-I", d]
    cxx = ["g++", "-std=c++17", "-O2", "-fPIC", "-pthread", "-w", "-include", os.path.join(HERE, "modem_abi.h")] + inc
    objs = []
    for src in REF_VISUAL_SRCS:                                  # compiled where they lie, objects only into oracle/_ref/
        o = os.path.join(vdir, src[:-4] + ".o")
        run(cxx + ["-c", os.path.join(proc, src), "-o", o])
        objs.append(o)
    drv = os.path.join(vdir, "ref_visuals.o")
    run(cxx + ["-c", os.path.join(HERE, "ref_visuals.cpp"), "-o", drv])
    abi = os.path.join(vdir, "ref_modem_abi.o")
    run(["gcc", "-O2", "-fPIC", "-c", os.path.join(HERE, "ref_modem_abi.c"), "-o", abi])
    run(["g++", "-shared", "-o", VISUALS_LIB, drv, abi] + objs +
        ["-L", OUT, "-l:libsdr_ref.so", "-Wl,-rpath,$ORIGIN", "-lm", "-pthread", "-Wl,--no-undefined"])
    return VISUALS_LIB


def build_demod(force=False):
    """oracle/_ref/libdemod_ref.so: the reference's unmodified DemodulatorInstance (pre-demod, worker and demodulator
    threads, modems) over its own liquid binary."""
    if not os.path.exists(REF_SRC):
        if os.path.exists(DEMOD_LIB):
            return DEMOD_LIB
        raise SystemExit("oracle/_ref/libdemod_ref.so missing and /root/reference absent")
    lib = build(force)
    shadow = os.path.join(HERE, "shadow")
    mine = [os.path.join(HERE, f) for f in ("modem_abi.h", "ref_modem_abi.c", "ref_demod.cpp", "ref_audio.cpp")] + \
           [os.path.join(shadow, "CubicSDR.h"), os.path.join(shadow, "AppConfig.h"), os.path.join(shadow, "SoapySDR", "Types.hpp"),
            os.path.abspath(__file__)]
    if (not force) and os.path.exists(DEMOD_LIB) and all(os.path.getmtime(DEMOD_LIB) >= os.path.getmtime(s) for s in mine + [lib]):
        return DEMOD_LIB
    ddir = os.path.join(OUT, "demod_obj")
    os.makedirs(ddir, exist_ok=True)
    mod = os.path.join(REF_SRC, "modules", "modem")
    inc = []
    for d in (shadow, REF_SRC, os.path.join(REF_SRC, "util"), os.path.join(REF_SRC, "audio"), os.path.join(REF_SRC, "demod"),
              os.path.join(REF_SRC, "sdr"), os.path.join(REF_SRC, "process"), mod, os.path.join(mod, "analog"), os.path.join(mod, "digital"),
              "/root/reference/external/liquid-dsp/include", "/root/reference/external/rtaudio"):
        inc += ["-I", d]
    cxx = ["g++", "-std=c++17", "-O2", "-fPIC", "-pthread", "-w", "-include", os.path.join(HERE, "modem_abi.h")] + inc
    objs = []
    srcs = [os.path.join(REF_SRC, f) for f in REF_DEMOD_SRCS] + [os.path.join(mod, f) for f in REF_MODEM_SRCS] + \
           ["/root/reference/external/rtaudio/RtAudio.cpp"]
    for src in srcs:                                             # compiled where they lie, objects only into oracle/_ref/
        o = os.path.join(ddir, os.path.basename(src)[:-4] + ".o")
        run(cxx + ["-c", src, "-o", o])
        objs.append(o)
    # src/audio/AudioThread.cpp enters oracle/ref_audio.cpp's translation unit by #include (its RtAudio callback is file-static)
    aud = os.path.join(ddir, "ref_audio.o")
    run(cxx + ["-fno-access-control", "-c", os.path.join(HERE, "ref_audio.cpp"), "-o", aud])
    objs.append(aud)
    drv = os.path.join(ddir, "ref_demod.o")
    run(cxx + ["-fno-access-control", "-c", os.path.join(HERE, "ref_demod.cpp"), "-o", drv])
    abi = os.path.join(ddir, "ref_modem_abi.o")
    run(["gcc", "-O2", "-fPIC", "-c", os.path.join(HERE, "ref_modem_abi.c"), "-o", abi])
    run(["g++", "-shared", "-o", DEMOD_LIB, drv, abi] + objs +
        ["-L", OUT, "-l:libsdr_ref.so", "-Wl,-rpath,$ORIGIN", "-lm", "-pthread", "-Wl,--no-undefined"])
    return DEMOD_LIB


if __name__ == "__main__":
    if "--clean" in sys.argv and os.path.exists(REF_AR) and os.path.isdir(OBJ):
        import shutil
        shutil.rmtree(OBJ)      # our own build products only; re-extracted from the reference archive below
    print(build(force="--force" in sys.argv or "--clean" in sys.argv))
    print(build_modems(force="--force" in sys.argv or "--clean" in sys.argv))
    print(build_visuals(force="--force" in sys.argv or "--clean" in sys.argv))
    print(build_demod(force="--force" in sys.argv or "--clean" in sys.argv))
