"""In-tree build of the CUDA library (sm_100a only; nvcc cross-compiles without a GPU)."""
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libimfeat.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    hdr = os.path.join(os.path.dirname(PKG_DIR), "include", "imfeat.h")
    return any(os.path.getmtime(s) > t for s in sources() + [hdr])


def build_library(force=False, verbose=False):
    """Compile csrc/imfeat_api.cu (which includes every kernel header) into libimfeat.so."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build %s" % LIB_PATH)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
        "-o", LIB_PATH, os.path.join(CSRC, "imfeat_api.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr))
    if verbose:
        print(res.stderr)
    return LIB_PATH


# ---- the thin PyTorch C++ extension over the C ABI (torch ops imfeat::extract / imfeat::glcm_counts) ----
TORCH_EXT_SRC = os.path.join(PKG_DIR, "csrc_torch", "imfeat_torch.cpp")
TORCH_EXT_PATH = os.path.join(PKG_DIR, "libimfeat_torch.so")


def torch_ext_needs_build():
    if not os.path.exists(TORCH_EXT_PATH):
        return True
    t = os.path.getmtime(TORCH_EXT_PATH)
    hdr = os.path.join(os.path.dirname(PKG_DIR), "include", "imfeat.h")
    return any(os.path.getmtime(f) > t for f in (TORCH_EXT_SRC, hdr))


def build_torch_extension(force=False):
    """g++ only (the file has no device code): ATen / c10 headers from the installed torch, linked against
    libimfeat.so next to it (rpath $ORIGIN) and torch's own libraries."""
    build_library()
    if not force and not torch_ext_needs_build():
        return TORCH_EXT_PATH
    import torch
    from torch.utils import cpp_extension as ce
    inc = ce.include_paths("cuda") if hasattr(ce, "include_paths") else []
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    gxx = shutil.which("g++") or "g++"
    cmd = [gxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI),
           "-DTORCH_API_INCLUDE_EXTENSION_H"]
    for d in inc:
        cmd += ["-isystem", d]
    cmd += ["-o", TORCH_EXT_PATH, TORCH_EXT_SRC, "-L", PKG_DIR, "-limfeat", "-L", tlib, "-lc10", "-lc10_cuda", "-ltorch_cpu",
            "-ltorch_cuda", "-ltorch", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + tlib]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building the torch extension failed:\n%s\n%s" % (" ".join(cmd), res.stderr[-4000:]))
    return TORCH_EXT_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
    print(build_torch_extension(force=True))
