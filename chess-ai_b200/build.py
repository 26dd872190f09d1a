"""Builds libchessai_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot. nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libchessai_b200.so")
SOURCES = ["cab_api.cu", "cab_train.cu", "cab_tc.cu", "cab_leafq.cu", "cab_prediction.cu", "cab_mcts.cu"]
HEADERS = ["cab_internal.cuh", "chess_rules.cuh", "encode_simd.cuh", "mlp_stream_kernel.cuh", "tc_gemm_bf16.cuh", "mlp_tc_fused.cuh", "train_kernels.cuh", os.path.join("..", "..", "include", "chessai_b200.h"),
           os.path.join("..", "host", "fastchess.hpp")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2",
    "--expt-relaxed-constexpr",
]
# private static C++ runtime: the library must not interpose libstdc++ symbols of the host process (python/torch)
LINK = ["-shared", "-cudart", "shared", "-Xlinker", "--exclude-libs=ALL", "-Xlinker", "-Bsymbolic",
        "-Xcompiler", "-static-libstdc++,-static-libgcc", "-ldl"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=False, force=False):
    os.makedirs(LIB_DIR, exist_ok=True)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    objs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(LIB_DIR, s.replace(".cu", ".o"))
        if force or _stale(obj, deps):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            subprocess.check_call(cmd)
        objs.append(obj)
    if force or _stale(LIB, objs):
        subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a"] + LINK + objs + ["-o", LIB])
    # host-side C++ (g++): the fast rules engine (C entry points for the tests) and the self-play driver, the trainer
    host = os.path.join(HERE, "host")
    gxx = os.environ.get("CXX", "g++")
    fc = os.path.join(LIB_DIR, "libfastchess.so")
    fc_deps = [os.path.join(host, "fastchess_c.cpp"), os.path.join(host, "fastchess.hpp"), os.path.join(host, "cases.hpp"),
               os.path.join(CSRC, "chess_rules.cuh")]
    if force or _stale(fc, fc_deps):
        subprocess.check_call([gxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-fvisibility=hidden", "-static-libstdc++",
                               "-static-libgcc", "-o", fc, fc_deps[0]])
    sp = os.path.join(LIB_DIR, "selfplay")
    sp_deps = [os.path.join(host, "selfplay.cpp"), os.path.join(host, "fastchess.hpp"), os.path.join(host, "cases.hpp"), LIB]
    if force or _stale(sp, sp_deps):
        subprocess.check_call([gxx, "-std=c++17", "-O2", "-pthread", "-static-libstdc++", "-static-libgcc", "-o", sp, sp_deps[0],
                               "-L" + LIB_DIR, "-lchessai_b200", "-Wl,-rpath,$ORIGIN"])
    tr = os.path.join(LIB_DIR, "train")
    tr_deps = [os.path.join(host, "train.cpp"), os.path.join(host, "cases.hpp"), LIB]
    if force or _stale(tr, tr_deps):
        subprocess.check_call([gxx, "-std=c++17", "-O2", "-static-libstdc++", "-static-libgcc", "-o", tr, tr_deps[0],
                               "-L" + LIB_DIR, "-lchessai_b200", "-Wl,-rpath,$ORIGIN"])
    fb = os.path.join(LIB_DIR, "fwd_bench")      # native timing harness of the forward path (kernel-variant sweeps)
    fb_src = os.path.join(CSRC, "tools", "fwd_bench.cpp")
    if force or _stale(fb, [fb_src, LIB]):
        subprocess.check_call([gxx, "-std=c++17", "-O2", "-static-libstdc++", "-static-libgcc", "-I/usr/local/cuda/include", "-o", fb, fb_src,
                               "-L" + LIB_DIR, "-lchessai_b200", "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath,$ORIGIN"])
    for name in ("pipe_peaks", "tma_stream", "act_bench", "mufu_bench", "tmem_bench", "mma_bench"):   # microbenchmarks behind the roofline denominators / design choices
        tool = os.path.join(LIB_DIR, name)
        tsrc = os.path.join(CSRC, "tools", name + ".cu")
        if force or _stale(tool, [tsrc]):
            subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-w", tsrc, "-o", tool])
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
