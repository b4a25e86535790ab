"""Stage the reference's own Python test-suite next to this repository's tests.

The acceptance test of a drop-in is the reference's test-suite run against it
(SURVEY.md section 2.3 / 7 step 1).  The reference's sources are never committed
here, so -- exactly like ``oracle/_ref`` (the reference core compiled from the
sources where they lie) -- ``/root/reference/tests/test_*.py`` are copied, unmodified,
into ``tests/zz_reference_suite/_staged/`` by this script (``__graft_entry__.build()``
runs it wherever the reference is mounted).  ``_staged/`` is git-ignored but not
gpurun-ignored: it travels to the GPU box with the built libraries, where
``pytest tests -m gpu`` collects it (``conftest.py`` in this directory marks every
staged test ``gpu`` and puts ``import adrt`` = this engine on the path).

    python tests/zz_reference_suite/stage.py [/root/reference]
"""
import glob
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_staged")


def stage(ref_root="/root/reference"):
    src = os.path.join(ref_root, "tests")
    files = sorted(glob.glob(os.path.join(src, "test_*.py")))
    if not files:
        print(f"reference tests not present at {src}; keeping what is staged")
        return 0
    os.makedirs(STAGED, exist_ok=True)
    for old in glob.glob(os.path.join(STAGED, "test_*.py")):
        os.remove(old)
    manifest = {}
    for f in files:
        dst = os.path.join(STAGED, os.path.basename(f))
        shutil.copyfile(f, dst)
        with open(dst, "rb") as fh:
            manifest[os.path.basename(f)] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(STAGED, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "files": manifest}, fh, indent=1, sort_keys=True)
    print(f"staged {len(files)} reference test modules into {STAGED}")
    return len(files)


if __name__ == "__main__":
    stage(*sys.argv[1:2])
