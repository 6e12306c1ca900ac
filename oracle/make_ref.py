"""ORACLE recipe (test infrastructure): make the UNMODIFIED reference modules of the hot path available
to the GPU box.  The reference is pure Python; its six hot-path modules are copied byte for byte from
the read-only reference tree into oracle/_ref/ (git-ignored: never part of this repository's history,
but not gpurun-ignored, so the directory travels with the snapshot like a built .so).  bench.py's CPU
arm (`--impl reference`, `cpu_baseline`) then times the reference's OWN classes (kind "reference");
without oracle/_ref it falls back to the oracle port (kind "port").

    python oracle/make_ref.py          # run in the container that has /root/reference
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("CLSKD_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("config.py", "tools_for_model.py", "tools_for_loss.py", "DCCRN.py", "feature_extraction.py", "framework.py")


def main():
    if not os.path.exists(os.path.join(SRC, "DCCRN.py")):
        print("oracle/make_ref.py: %s not found - nothing to do (the GPU box uses the prebuilt oracle/_ref)" % SRC)
        return 0
    os.makedirs(DST, exist_ok=True)
    manifest = []
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        manifest.append("%s  %s" % (hashlib.sha256(open(os.path.join(DST, f), "rb").read()).hexdigest(), f))
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as fh:
        fh.write("\n".join(manifest) + "\n")
    print("oracle/_ref: %d reference modules copied unmodified from %s" % (len(FILES), SRC))
    return 0


if __name__ == "__main__":
    sys.exit(main())
