"""Build the CUDA extension in-tree: lives_b200/libpe_b200.so (sm_100a only).

`python -m lives_b200.build` or __graft_entry__.build().  nvcc cross-compiles without a GPU.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpe_b200.so")
PLUGIN = os.path.join(HERE, "libpe_weed_plugin.so")

SOURCES = ["pe_engine.cu", "pe_kernels_rgb.cu", "pe_kernels_yuv.cu", "pe_kernels_yuv2.cu", "pe_kernels_yuv3.cu", "pe_kernels_fused.cu", "pe_kernels_fused2.cu", "pe_kernels_fused3.cu", "pe_tables.cpp"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "pixel_engine.h")]
    deps = [d for d in deps if os.path.isfile(d)]
    if force or _stale(LIB, deps):
        extra = ["-DPE_F3_NT=" + os.environ["PE_F3_NT"]] if os.environ.get("PE_F3_NT") else []  # tuning experiments only
        extra += os.environ.get("PE_NVCC_EXTRA", "").split()
        cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd, cwd=CSRC)
    plugin_src = os.path.join(CSRC, "pe_weed_plugin.c")
    if os.path.exists(plugin_src) and (force or _stale(PLUGIN, [plugin_src, LIB] + deps)):
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=gnu11", "-Wall", "-I", os.path.join(HERE, "..", "include"),
               "-o", PLUGIN, plugin_src, "-L", HERE, "-lpe_b200", "-Wl,-rpath,$ORIGIN", "-ldl"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB)
