"""Ship the UNMODIFIED reference to the GPU box as ``oracle/_ref`` (TEST INFRASTRUCTURE ONLY).

``/root/reference`` exists only in the build container.  This recipe copies the reference's own Python package
(``cerberusdet/``: 40 ``.py`` + 4 ``.yaml`` files, 476 KB) and its ``data/*.yaml`` / ``data/hyps`` configs, file for file,
into ``oracle/_ref/`` -- git-ignored (no reference source enters the history) but not gpurun-ignored, so it travels with
the snapshot like the built ``.so`` files.  On the GPU box the ``-m gpu`` drop-in tests, ``bench.py``'s same-GPU eager
baseline and ``--impl reference`` then run the reference ITSELF (``oracle/ref_import.py`` falls back to this copy when
``/root/reference`` is absent); nothing under ``cerberusdet_b200/`` ever imports it.

    python oracle/make_ref.py            # called by __graft_entry__.build() when /root/reference is present
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("CERB_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")


def make_ref(force: bool = False) -> str | None:
    if not os.path.isdir(os.path.join(SRC, "cerberusdet")):
        return DST if os.path.isdir(os.path.join(DST, "cerberusdet")) else None
    if force and os.path.isdir(DST):
        shutil.rmtree(DST)
    keep = (".py", ".yaml", ".yml", ".toml", ".txt", ".md")
    for sub in ("cerberusdet", "data"):
        for root, dirs, files in os.walk(os.path.join(SRC, sub)):
            dirs[:] = [d for d in dirs if d not in ("__pycache__", "images", "scripts")]
            rel = os.path.relpath(root, SRC)
            os.makedirs(os.path.join(DST, rel), exist_ok=True)
            for f in files:
                if f.endswith(keep):
                    shutil.copy2(os.path.join(root, f), os.path.join(DST, rel, f))
    for f in ("LICENSE.txt", "pyproject.toml"):
        if os.path.exists(os.path.join(SRC, f)):
            shutil.copy2(os.path.join(SRC, f), os.path.join(DST, f))
    return DST


if __name__ == "__main__":
    out = make_ref(force="--force" in sys.argv)
    print("oracle/_ref:", out)
