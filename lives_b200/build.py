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
LAYERLIB = os.path.join(HERE, "libpe_weed_layer.so")
VPPLIB = os.path.join(HERE, "libpe_vpp.so")

SOURCES = ["pe_engine.cu", "pe_kernels_rgb.cu", "pe_kernels_yuv.cu", "pe_kernels_yuv2.cu", "pe_kernels_yuv3.cu", "pe_kernels_fused.cu", "pe_kernels_fused2.cu", "pe_kernels_fused3.cu", "pe_kernels_fused4.cu", "pe_kernels_fx2.cu", "pe_kernels_float.cu", "pe_kernels_mc.cu", "pe_tables.cpp", "pe_hoststage.cpp"]
OBJ = os.path.join(HERE, "build")
NVCC_COMPILE = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC"]
NVCC_LINK = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "shared"]


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


def _compile_one(args):
    src, obj, extra, verbose = args
    cmd = [_nvcc()] + NVCC_COMPILE + extra + ["-c", "-o", obj, src]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd, cwd=CSRC)


def build(force=False, verbose=False):
    """every source to its own object (in parallel, only the stale ones), then one link: iterating on one kernel file costs one
    nvcc run"""
    from concurrent.futures import ThreadPoolExecutor
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers += [os.path.join(HERE, "..", "include", f) for f in os.listdir(os.path.join(HERE, "..", "include"))]
    os.makedirs(OBJ, exist_ok=True)
    extra = ["-DPE_F3_NT=" + os.environ["PE_F3_NT"]] if os.environ.get("PE_F3_NT") else []  # tuning experiments only
    extra += os.environ.get("PE_NVCC_EXTRA", "").split()
    force = force or bool(extra)
    jobs, objs = [], []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, os.path.splitext(s)[0] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            jobs.append((src, obj, extra, verbose))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(_compile_one, jobs))
    if jobs or not os.path.exists(LIB):
        cmd = [_nvcc()] + NVCC_LINK + ["-o", LIB] + objs
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd, cwd=CSRC)
    deps = headers
    plugin_src = os.path.join(CSRC, "pe_weed_plugin.c")
    if os.path.exists(plugin_src) and (force or _stale(PLUGIN, [plugin_src, LIB] + deps)):
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=gnu11", "-Wall", "-I", os.path.join(HERE, "..", "include"),
               "-o", PLUGIN, plugin_src, "-L", HERE, "-lpe_b200", "-Wl,-rpath,$ORIGIN", "-ldl"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd, cwd=CSRC)
    layer_src = os.path.join(CSRC, "pe_weed_layer.c")
    if os.path.exists(layer_src) and (force or _stale(LAYERLIB, [layer_src, LIB] + deps)):
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=gnu11", "-Wall", "-I", os.path.join(HERE, "..", "include"),
               "-o", LAYERLIB, layer_src, "-L", HERE, "-lpe_b200", "-Wl,-rpath,$ORIGIN", "-ldl"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd, cwd=CSRC)
    vpp_src = os.path.join(CSRC, "pe_vpp.c")
    if os.path.exists(vpp_src) and (force or _stale(VPPLIB, [vpp_src, LIB] + deps)):
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=gnu11", "-Wall", "-I", os.path.join(HERE, "..", "include"),
               "-o", VPPLIB, vpp_src, "-L", HERE, "-lpe_b200", "-Wl,-rpath,$ORIGIN", "-ldl"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB)
