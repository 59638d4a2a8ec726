"""TEST INFRASTRUCTURE ONLY -- recipe that makes the UNMODIFIED reference travel to the GPU box.

    python -m oracle.build_ref          (also run by __graft_entry__.build() when /root/reference exists)

The reference is pure Python (a vendored `deepctr` package plus the xdftrain*.py caller scripts): there is nothing to compile, so
"building" it is a byte-for-byte copy of its Python sources from /root/reference into `oracle/_ref/` -- git-ignored (never part of
this repository's history), NOT gpurun-ignored (it ships with the snapshot, like the in-tree .so).  Consumers:
  * `bench.py --impl reference` / the `cpu_baseline` leg: times the reference's own train step on the box's host cores
    (`cpu_baseline.kind == "reference"`; without oracle/_ref the oracle port is timed and the line says "port");
  * `tests/test_gpu_dropin_script.py`: runs the unmodified `xdftrain.py` against THIS repo's `deepctr` package.
Nothing in the product package imports it.  A MANIFEST with sha256 of every copied file is written next to the copy so a reader can
check that the files are the reference's own.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("XDFM_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
KEEP_EXT = (".py", ".bash", ".txt", ".md")


def build_ref(force=False):
    """Copy the reference's sources; returns the destination or None when /root/reference is absent (GPU box: use what travelled)."""
    if not os.path.isdir(os.path.join(SRC, "deepctr")):
        return DST if os.path.isdir(os.path.join(DST, "deepctr")) else None
    manifest = {}
    for root, dirs, files in os.walk(SRC):
        dirs[:] = [d for d in dirs if d not in (".git", "__pycache__")]
        for f in files:
            if not (f.endswith(KEEP_EXT) or f == "LICENSE"):
                continue
            src = os.path.join(root, f)
            rel = os.path.relpath(src, SRC)
            dst = os.path.join(DST, rel)
            data = open(src, "rb").read()
            manifest[rel] = hashlib.sha256(data).hexdigest()
            if not force and os.path.exists(dst) and open(dst, "rb").read() == data:
                continue
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    return DST


if __name__ == "__main__":
    out = build_ref(force="--force" in sys.argv)
    print("reference copy:", out)
