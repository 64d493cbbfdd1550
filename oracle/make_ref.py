"""Stage the UNMODIFIED reference module for the CPU arm of bench.py.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference is one pure-Python file,
/root/reference/src/model.py; it exists in the build container only.  This script copies it, byte for
byte, to oracle/_ref/src/model.py -- a git-ignored directory (never part of the repository's history)
that still travels to the GPU box with the working tree, like the built .so files -- so that
``bench.py --impl reference`` and the ``cpu_baseline`` leg time the reference's own code on the box's
host cores (``kind: "reference"``) instead of the op-for-op port in oracle/torch_port.py.
``__graft_entry__.build()`` runs it whenever /root/reference is present.

    python -m oracle.make_ref        # prints the sha256 of the staged file
"""
from __future__ import annotations

import hashlib
import os
import shutil

REF_SRC = "/root/reference/src/model.py"
REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REF_DST = os.path.join(REF_DIR, "src", "model.py")


def stage() -> str | None:
    """Copy the reference module if the reference checkout is present; returns the staged path (or None)."""
    if not os.path.exists(REF_SRC):
        return REF_DST if os.path.exists(REF_DST) else None
    os.makedirs(os.path.dirname(REF_DST), exist_ok=True)
    shutil.copyfile(REF_SRC, REF_DST)
    with open(os.path.join(REF_DIR, "src", "__init__.py"), "w"):
        pass
    with open(os.path.join(REF_DIR, "SOURCE.txt"), "w") as f:
        f.write(f"byte-for-byte copy of {REF_SRC}\nsha256 {sha256(REF_DST)}\n")
    return REF_DST


def sha256(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def load():
    """The staged reference module (``LineRefineNet`` etc.), or None when it has not been staged."""
    if not os.path.exists(REF_DST):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("_lrn_reference_model", REF_DST)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = stage()
    print(p, sha256(p) if p else "reference checkout not present")
