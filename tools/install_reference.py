"""Put an UNMODIFIED copy of the reference where the benchmark's CPU arm can import it on the GPU box.

    python tools/install_reference.py [/root/reference]

The reference (zeynelacikgoez/PyAudioLocalization) is six plain Python modules without a setup.py, so `pip install`
has nothing to build: this recipe copies `*.py` verbatim into `baseline/_ref/` (git-ignored, NOT gpurun-ignored: it
travels to the GPU box with the snapshot, where /root/reference does not exist) and records the SHA-256 of every file
in `baseline/_ref/MANIFEST.json`.  Nothing under baseline/_ref is product source; bench.py imports it only in the
`--impl reference` / `cpu_baseline` legs, behind stub modules for the three absent third-party packages the hot path
never touches (soundfile, resampy, matplotlib).
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")


def install(src="/root/reference"):
    if not os.path.isdir(src):
        return None
    os.makedirs(DEST, exist_ok=True)
    manifest = {}
    for f in sorted(os.listdir(src)):
        if f.endswith(".py") or f in ("LICENSE", "requirement.txt"):
            shutil.copyfile(os.path.join(src, f), os.path.join(DEST, f))
            manifest[f] = hashlib.sha256(open(os.path.join(DEST, f), "rb").read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "files": manifest}, fh, indent=1)
    return DEST


if __name__ == "__main__":
    print(install(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))
