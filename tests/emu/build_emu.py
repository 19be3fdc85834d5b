"""Build tests/emu/_build/libpal_emu.so (host emulation of the CUDA kernel bodies; test only)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "pyaudiolocalization_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libpal_emu.so")


def build(force=False):
    srcs = [os.path.join(HERE, "emu_pal.cpp")] + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))
                                                   if f.endswith((".h", ".cuh"))]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in srcs):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-DPAL_EMU", "-fno-strict-aliasing", "-x", "c++", "-I", CSRC,
           os.path.join(HERE, "emu_pal.cpp"), "-o", OUT, "-lpthread"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
