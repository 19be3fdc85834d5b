"""Build libpal_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m pyaudiolocalization_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
ROOT = os.path.dirname(PKG)
LIB = os.environ.get("PAL_B200_LIB") or os.path.join(PKG, "libpal_b200.so")

NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def sources():
    out = [os.path.join(ROOT, "include", "pal_b200.h")]
    out += [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".h", ".cuh", ".cu"))]
    return out


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in sources()):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("PAL_NVCC_EXTRA", "").split()      # tuning experiments only
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-I", CSRC, "-I", os.path.join(ROOT, "include"),
                                 os.path.join(CSRC, "pal_capi.cu"), "-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(PKG, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if verbose or res.returncode != 0:
        print(log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed; see " + os.path.join(PKG, "build.log"))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
