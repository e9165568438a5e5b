"""Recipe for oracle/_ref (TEST / BASELINE INFRASTRUCTURE — never imported by the product path).

The reference is a plain Python package, so "building" it means placing an UNMODIFIED copy of
/root/reference/package/whisper-at/whisper_at under oracle/_ref/whisper_at.  oracle/_ref/ is git-ignored (reference
sources never enter this repository's history) but not gpurun-ignored, so the copy travels to the GPU box, where
/root/reference does not exist, and `bench.py --impl reference` / the `cpu_baseline` leg can time the reference ITSELF
(`kind: "reference"`) instead of the oracle port.  Run here by `__graft_entry__.build()` whenever /root/reference is
present; a box without oracle/_ref falls back to the port and says so.

    python oracle/make_ref.py            # copy + write oracle/_ref/PROVENANCE.json (sha256 of every file)
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/package/whisper-at/whisper_at"
DST_ROOT = os.path.join(ROOT, "oracle", "_ref")
DST = os.path.join(DST_ROOT, "whisper_at")


def _make_writable(root: str) -> None:
    for base, dirs, files in os.walk(root):
        os.chmod(base, 0o755)
        for f in files:
            os.chmod(os.path.join(base, f), 0o644)


def make_ref(verbose: bool = True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"make_ref: {SRC} not present; keeping whatever is in {DST_ROOT}")
        return os.path.isdir(DST)
    if os.path.isdir(DST):
        _make_writable(DST)
        shutil.rmtree(DST)
    os.makedirs(DST_ROOT, exist_ok=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"), copy_function=shutil.copyfile)
    _make_writable(DST)                                  # the source tree is read-only; the copy must stay replaceable
    prov = {}
    for base, _, files in os.walk(DST):
        for f in sorted(files):
            p = os.path.join(base, f)
            with open(p, "rb") as fh:
                prov[os.path.relpath(p, DST_ROOT)] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DST_ROOT, "PROVENANCE.json"), "w") as fh:
        json.dump(dict(source=SRC, note="verbatim copy made by oracle/make_ref.py; not tracked by git", sha256=prov), fh, indent=1)
    if verbose:
        print(f"make_ref: copied {len(prov)} files to {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make_ref() else 1)
