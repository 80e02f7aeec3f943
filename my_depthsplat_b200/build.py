"""Build libb200splat.so in-tree with nvcc for sm_100a (``python -m my_depthsplat_b200.build``)."""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"


def build(force: bool = False, verbose: bool = False, extra: str = "") -> Path:
    if force:
        subprocess.run(["make", "-C", str(CSRC), "clean"], check=True, capture_output=not verbose)
    cmd = ["make", "-C", str(CSRC), "-j8"] + ([f"EXTRA={extra}"] if extra else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("building libb200splat.so failed")
    return CSRC.parent / "libb200splat.so"


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
