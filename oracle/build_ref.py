#!/usr/bin/env python
"""Recipe for oracle/_ref/: a copy of the UNMODIFIED reference's simulator packages, so that the reference's own
single-process CPU implementation can be timed on the GPU box beside the CUDA path (bench.py: cpu_baseline
"reference_python_1core").

    python oracle/build_ref.py            # needs /root/reference (build container only)

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  oracle/_ref/ is git-ignored (no reference source enters the history) but
not gpurun-ignored, so it travels to the GPU box with the snapshot like the built .so files; /root/reference itself
does not exist there.  Nothing is edited: the files are copied byte for byte from where they lie under
/root/reference (rl_env/, run_colav/, utils/: Python sources and the route tables).  The packages the reference
imports but the image lacks (matplotlib, gymnasium, shapely, gtimer) are provided at import time by the stub modules
of oracle/ref_harness.py -- shapely by the documented-semantics stand-in pinned in tests/test_map_geometry.py.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("AST_SAC_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
PACKAGES = ("rl_env", "run_colav", "utils")
KEEP = (".py", ".txt")


def build() -> str:
    if not os.path.isdir(os.path.join(SRC, "run_colav")):
        raise SystemExit(f"{SRC} is not the reference tree")
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for pkg in PACKAGES:
        for root, _, files in os.walk(os.path.join(SRC, pkg)):
            for f in files:
                if not f.endswith(KEEP):
                    continue
                src = os.path.join(root, f)
                rel = os.path.relpath(src, SRC)
                dst = os.path.join(DST, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
                manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    return DST


if __name__ == "__main__":
    out = build()
    n = sum(len(f) for _, _, f in os.walk(out))
    print(f"{out}: {n} files")
    sys.exit(0)
