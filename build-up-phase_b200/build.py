"""Build recipe of librtcore.so (CUDA, sm_100a only). nvcc cross-compiles without a GPU.

-fmad=false and -ffp-contract=off are part of the numerical contract (bit parity of hit ids with the
strict-IEEE CPU oracle), not tuning knobs. -lineinfo keeps ncu's source page mapped to our code.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librtcore.so")
SOURCES = ["rtcore_api.cu", "rt_group.cu", "lbvh_build.cu", "radix_sort.cu", "trace.cu",
           os.path.join("..", "host", "rtcore_io.cpp")]     # host-side .obj / image I/O (include/rtcore_io.h), no device code
HEADERS = ["rt_internal.h", "rt_host.h", "rt_device.cuh", "seg_sort.cuh", os.path.join("..", "..", "include", "rtcore.h"),
           os.path.join("..", "..", "include", "rtcore_io.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden,-O2",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    files = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(f) > t for f in files)


def build_variant(name: str, defines) -> str:
    """Tuning aid: builds build/librtcore_<name>.so with extra -D flags (select it with RTCORE_LIB=<path>)."""
    out = os.path.join(HERE, "build", f"librtcore_{name}.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    flags = [f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")]
    cmd = [_nvcc(), "-ccbin", ccbin, *flags, *[f"-D{d}" for d in defines], "-shared", "-o", out,
           *[os.path.join(CSRC, s) for s in SOURCES], "-lcudart", "-lrt"]
    subprocess.check_call(cmd)
    return out


def build_rtcore(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(objdir, os.path.splitext(os.path.basename(s))[0] + ".o")
        objs.append(o)
        cmd = [_nvcc(), "-ccbin", ccbin, *NVCC_FLAGS, "-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for s, p in procs:
        out, _ = p.communicate()
        log.append(f"== {s}\n{out}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {s}")
    cmd = [_nvcc(), "-ccbin", ccbin, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *objs, "-lcudart", "-lrt"]
    subprocess.check_call(cmd)
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


def build_host_sample(name: str = "sample_scene") -> str:
    """Compiles a headless C++ host program (host/sample_scene.cpp: the reference's main(); host/sample_scene_mgpu.cpp: the same on
    the GPUs of one box, one process per GPU) against librtcore.so."""
    src = os.path.join(HERE, "host", name + ".cpp")
    out = os.path.join(HERE, "host", name)
    if os.path.exists(out) and os.path.getmtime(out) > max(os.path.getmtime(src), os.path.getmtime(LIB)):
        return out
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", src, "-I", os.path.join(HERE, "..", "include"), "-L", HERE, "-lrtcore",
                           "-Wl,-rpath," + HERE, "-Wl,-rpath,/usr/local/cuda/lib64", "-o", out])
    return out


if __name__ == "__main__":
    print(build_rtcore(force="--force" in sys.argv, verbose=True))
    print(build_host_sample())
    print(build_host_sample("sample_scene_mgpu"))
