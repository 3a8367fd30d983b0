"""TEST INFRASTRUCTURE ONLY. Picks the build of the unmodified reference (oracle/Makefile -> oracle/_ref/) that fits
the host CPU: the x86-64-v4 (AVX-512) binaries when /proc/cpuinfo lists the level's features, else x86-64-v3.
(-march=native, the reference's own flag, cannot travel from the build container to the GPU box.)"""
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_V4 = ("avx512f", "avx512bw", "avx512vl", "avx512dq", "avx512cd")


def cpu_level():
    try:
        flags = set()
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("flags"):
                flags = set(ln.split(":", 1)[1].split())
                break
        return "x86-64-v4" if all(f in flags for f in _V4) else "x86-64-v3"
    except OSError:
        return "x86-64-v3"


def ref_path(name):
    """name: 'grlbwt_ref' (the reference CLI) or 'ref_harness' (its par_phase alone + per-round dumps); '' if not built"""
    if cpu_level() == "x86-64-v4":
        p = os.path.join(HERE, "_ref", name + "_v4")
        if os.path.exists(p):
            return p
    p = os.path.join(HERE, "_ref", name)
    return p if os.path.exists(p) else ""


def ref_flags():
    return "-O3 -funroll-loops -fomit-frame-pointer -ffast-math -msse4.2 -march=" + cpu_level() + " (reference CMake flags; -march=native replaced by the portable level), asserts on"
