"""Builds libalacgpu.so (CUDA kernels + C ABI) and libalacnet_host.so (C++ host
mirror of AlacContext / QtMovieT / ALACFileReader) in-tree for sm_100a."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB_GPU = os.path.join(HERE, "libalacgpu.so")
LIB_GPU_CHECKED = os.path.join(HERE, "libalacgpu_checked.so")   # -DALACGPU_CHECKED: bounds assertions in every kernel
LIB_HOST = os.path.join(HERE, "libalacnet_host.so")

CU_SOURCES = ["k0_index.cu", "k12_decode.cu", "k3_stereo.cu", "kf_frame.cu", "runtime.cu"]
OBJ_DIR = os.path.join(ROOT, "build", "obj")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall",
]
NVCC_LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-cudart", "static"]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_gpu(force: bool = False, verbose: bool = False, checked: bool = False) -> str:
    """libalacgpu.so, or (checked) libalacgpu_checked.so: the same sources with -DALACGPU_CHECKED, the stand-in for
    compute-sanitizer (closed on this GPU pool) that the -m gpu tests run the fuzz / malformed-stream cases under."""
    LIB_GPU = LIB_GPU_CHECKED if checked else globals()["LIB_GPU"]
    OBJ_DIR = globals()["OBJ_DIR"] + ("_checked" if checked else "")
    srcs = [os.path.join(CSRC, s) for s in CU_SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".h", ".cuh"))]
    deps.append(os.path.join(ROOT, "include", "alacgpu.h"))
    if force or _newer(LIB_GPU, deps):
        # one object per source, compiled side by side (the frame-lane kernels alone take ~40 s), then one link
        from concurrent.futures import ThreadPoolExecutor
        os.makedirs(OBJ_DIR, exist_ok=True)
        extra = ["-DALACGPU_CHECKED"] if checked else []

        def compile_one(src: str) -> str:
            obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
            if force or _newer(obj, [src] + [d for d in deps if not d.endswith(".cu")]):
                cmd = ["nvcc", *NVCC_FLAGS, *extra, "-I", os.path.join(ROOT, "include"), "-c", "-o", obj, src]
                if verbose:
                    cmd.insert(1, "-Xptxas=-v")
                    print(" ".join(cmd), file=sys.stderr)
                subprocess.check_call(cmd)
            return obj

        with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
            objs = list(ex.map(compile_one, srcs))
        subprocess.check_call(["nvcc", *NVCC_LINK_FLAGS, "-o", LIB_GPU, *objs])
    return LIB_GPU


def build_host(force: bool = False) -> str:
    if not os.path.isdir(HOST):
        return ""
    srcs = [os.path.join(HOST, s) for s in sorted(os.listdir(HOST)) if s.endswith(".cpp")]
    if not srcs:
        return ""
    deps = srcs + [os.path.join(HOST, h) for h in os.listdir(HOST) if h.endswith((".h", ".hpp"))]
    deps.append(os.path.join(ROOT, "include", "alacgpu.h"))
    if force or _newer(LIB_HOST, deps):
        # links against libalacgpu.so next to it (rpath $ORIGIN): the mirror has no decode code of its own
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-fvisibility=hidden",
               "-I", os.path.join(ROOT, "include"), "-o", LIB_HOST, *srcs,
               "-L", HERE, "-lalacgpu", "-Wl,-rpath,$ORIGIN", "-ldl", "-lpthread"]
        subprocess.check_call(cmd)
    return LIB_HOST


CLI = os.path.join(HERE, "cli")
BIN_CLI = os.path.join(HERE, "alacgpu_decode")


def build_cli(force: bool = False) -> str:
    """alacgpu_decode: batch .m4a -> .wav tool over libalacgpu.so + the host mirror."""
    src = os.path.join(CLI, "alacgpu_decode.cpp")
    if not os.path.exists(src):
        return ""
    host_srcs = [os.path.join(HOST, s) for s in sorted(os.listdir(HOST)) if s.endswith(".cpp")]
    deps = [src, LIB_GPU, *host_srcs] + [os.path.join(HOST, h) for h in os.listdir(HOST) if h.endswith(".hpp")]
    if force or _newer(BIN_CLI, deps):
        # the host mirror is compiled in (its shared library only exports the flat C test surface)
        cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", BIN_CLI, src, *host_srcs,
               "-L", HERE, "-lalacgpu", "-Wl,-rpath,$ORIGIN", "-ldl", "-lpthread"]
        subprocess.check_call(cmd)
    return BIN_CLI


def build_all(force: bool = False, verbose: bool = False) -> None:
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=2) as ex:           # release and checked builds side by side
        list(ex.map(lambda c: build_gpu(force, verbose, checked=c), (False, True)))
    build_host(force)
    build_cli(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB_GPU)
