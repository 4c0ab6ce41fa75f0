"""TEST INFRASTRUCTURE -- recipe that stages the UNMODIFIED reference packages for the CPU reference arm.

    python oracle/make_ref.py            # copies /root/reference/{rdesign,rnampnn}/**/*.py -> oracle/_ref/

``oracle/_ref/`` is git-ignored (reference sources never enter this repository's history) but is not
gpurun-ignored, so the staged copy travels to the GPU box, where ``/root/reference`` does not exist.  It is used by
``bench.py --impl reference`` / ``cpu_baseline`` (kind "reference": the reference's own ``RNAModel`` on the host
cores) and by ``scripts/bench_reference_gpu.py`` (stock PyTorch on the B200).  Nothing on the product path imports it.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RNAMPNN_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
PACKAGES = ("rdesign", "rnampnn")


def stage(verbose: bool = True) -> bool:
    """Copy the reference's Python packages; returns False (and leaves ``_ref`` alone) if the reference is absent."""
    if not os.path.isdir(os.path.join(REF, "rdesign")):
        return False
    n = 0
    for pkg in PACKAGES:
        for root, _dirs, files in os.walk(os.path.join(REF, pkg)):
            rel = os.path.relpath(root, REF)
            if "__pycache__" in rel:
                continue
            os.makedirs(os.path.join(DST, rel), exist_ok=True)
            for f in files:
                if f.endswith(".py"):
                    shutil.copyfile(os.path.join(root, f), os.path.join(DST, rel, f))
                    n += 1
    with open(os.path.join(DST, "STAGED_FROM"), "w") as f:
        f.write(f"{REF}\n{n} files, unmodified\n")
    if verbose:
        print(f"[oracle/make_ref] staged {n} reference files under {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
