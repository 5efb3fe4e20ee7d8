"""In-tree build of the native code: libpcu_b200.so (CUDA kernels + C ABI, sm_100a only) and the
pybind11 module _pcu_internal that sits on top of it.

    python -m pcu_b200._build          # or:  import __graft_entry__; __graft_entry__.build()

nvcc cross-compiles for sm_100a without a GPU.  The artefacts stay next to the package (they are
git-ignored but travel to the GPU box with the gpurun snapshot).
"""
import os
import subprocess
import sys
import sysconfig

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
ROOT = os.path.dirname(PKG_DIR)
LIB = os.path.join(PKG_DIR, "libpcu_b200.so")
EXT = os.path.join(PKG_DIR, "_pcu_internal" + sysconfig.get_config_var("EXT_SUFFIX"))
HOST_CXX = "/usr/bin/g++"  # the image exports CXX=/opt/gcc/bin/g++, a wrapper that lacks parts of the toolchain

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--compiler-bindir", HOST_CXX,
    "-Xcompiler", "-fPIC",
    "-shared",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _sources(exts):
    out = []
    for base in (CSRC, os.path.join(ROOT, "include")):
        for f in sorted(os.listdir(base)):
            if f.endswith(exts):
                out.append(os.path.join(base, f))
    return out


def _run(cmd):
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("build step failed:\n  %s\n%s%s" % (" ".join(cmd), proc.stdout, proc.stderr))
    return proc.stdout + proc.stderr


def nvcc_path():
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    return cand if os.path.exists(cand) else "nvcc"


def build(force=False, verbose=False):
    """Compile whatever is out of date.  Returns the list of artefacts."""
    import pybind11

    cuda_src = _sources((".cu", ".cuh", ".inl", ".h"))
    if force or not _newer(LIB, cuda_src):
        cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              ["-o", LIB, os.path.join(CSRC, "pcu_b200.cu")]
        log = _run(cmd)
        if verbose:
            print(log)
    bind_src = [os.path.join(CSRC, "binding.cpp"), os.path.join(ROOT, "include", "pcu_b200.h")]
    if force or not _newer(EXT, bind_src + [LIB]):
        cmd = [HOST_CXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden",
               "-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"],
               os.path.join(CSRC, "binding.cpp"), "-o", EXT,
               "-L" + PKG_DIR, "-lpcu_b200", "-Wl,-rpath,$ORIGIN"]
        _run(cmd)
    return [LIB, EXT]


if __name__ == "__main__":
    for path in build(force="--force" in sys.argv, verbose="-v" in sys.argv):
        print(path)
