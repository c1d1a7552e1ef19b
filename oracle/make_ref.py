"""oracle/make_ref.py -- TEST / BASELINE INFRASTRUCTURE: recipe for `oracle/_ref/`.

The reference (squellww/super-resolution-system) is pure Python: there is nothing to compile, its "build" is putting
the three algorithm modules of the hot path on an import path.  This recipe copies them, UNMODIFIED, from where they
lie under /root/reference into the git-ignored directory `oracle/_ref/` (listed in .gitignore, not in .gpurunignore:
it travels to the GPU box like the built .so files and never enters the history).  `__graft_entry__.build()` runs it
whenever /root/reference is present (the build container); on the GPU box the prebuilt copy is used as is.

    python oracle/make_ref.py            # copy + import check
    python oracle/make_ref.py --check    # only report whether oracle/_ref is usable

Nothing under oracle/_ref is imported by the product package: only bench.py's `--impl reference` arm and the
cpu_baseline leg (through oracle/ref_arm.py) and tests/ touch it.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")
# the hot path's modules: tiler, Laplacian fusion, full-reference metrics
MODULES = ("tiling_module.py", "blending_module.py", "quality_assessment_module.py")


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DST, m)) for m in MODULES)


def make(verbose: bool = False) -> bool:
    """copy the modules if the reference tree is here; returns whether oracle/_ref is usable afterwards"""
    if not os.path.isdir(REF_SRC):
        return available()
    os.makedirs(REF_DST, exist_ok=True)
    manifest = {}
    for m in MODULES:
        src, dst = os.path.join(REF_SRC, m), os.path.join(REF_DST, m)
        if not os.path.exists(src):
            raise FileNotFoundError(f"{src}: the reference tree does not hold the hot path's module {m}")
        if not os.path.exists(dst) or _sha(src) != _sha(dst):
            shutil.copyfile(src, dst)
        manifest[m] = _sha(dst)
    with open(os.path.join(REF_DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_SRC, "sha256": manifest,
                   "note": "verbatim copies made by oracle/make_ref.py; git-ignored, never committed"}, f, indent=1)
    if verbose:
        print(f"oracle/_ref: {len(MODULES)} reference modules in place")
    return True


if __name__ == "__main__":
    if "--check" in sys.argv:
        print("oracle/_ref usable" if available() else "oracle/_ref missing")
        sys.exit(0 if available() else 1)
    ok = make(verbose=True)
    if ok:
        sys.path.insert(0, HERE + "/..")
        from oracle import ref_arm
        ref_arm.load()
        print("import check ok:", ", ".join(MODULES))
    sys.exit(0 if ok else 1)
