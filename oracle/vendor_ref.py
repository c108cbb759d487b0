"""Recipe that makes the UNMODIFIED reference hot-path sources available on the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

``/root/reference`` exists only in the build container.  ``bench.py --impl reference`` and ``cpu_baseline`` are run on the GPU box's
host cores, so the files of the reference that implement the hot path are copied — byte for byte, nothing edited — into
``oracle/_ref/`` by this script (called from ``__graft_entry__.build()`` whenever ``/root/reference`` is present).  ``oracle/_ref/`` is
listed in ``.gitignore`` (reference sources never enter this repository's history) but not in ``.gpurunignore``, so it travels to the GPU
box with the snapshot exactly like the built ``libtiab200.so``.  ``oracle/ref_shim.py`` then loads the modules from whichever root exists.

    python -m oracle.vendor_ref          # copy; prints the manifest
    python -m oracle.vendor_ref --check  # verify that oracle/_ref equals /root/reference for every listed file
"""
from __future__ import annotations

import hashlib
import json
import shutil
import sys
from pathlib import Path

SRC = Path("/root/reference")
DST = Path(__file__).resolve().parent / "_ref"

#: every file the shim reads or imports (paths relative to the reference root)
FILES = [
    "LICENSE",
    "tiatoolbox/__init__.py",
    "tiatoolbox/data/__init__.py",
    "tiatoolbox/data/pretrained_model.yaml",
    "tiatoolbox/data/remote_samples.yaml",
    "tiatoolbox/data/target_image.png",
    "tiatoolbox/utils/exceptions.py",
    "tiatoolbox/utils/misc.py",            # only the text of seven functions is used (AST-extracted by the shim)
    "tiatoolbox/utils/transforms.py",
    "tiatoolbox/tools/stainnorm.py",
    "tiatoolbox/tools/stainextract.py",
    "tiatoolbox/tools/stainaugment.py",
    "tiatoolbox/tools/patchextraction.py",  # text only: PatchExtractor.get_coordinates
    "tiatoolbox/models/models_abc.py",
    "tiatoolbox/models/architecture/utils.py",
    "tiatoolbox/models/architecture/hovernet.py",
    "tiatoolbox/models/architecture/hovernetplus.py",
    "tiatoolbox/models/architecture/unet.py",
    "tiatoolbox/models/architecture/vanilla.py",
    "tiatoolbox/models/engine/semantic_segmentor.py",    # text only: merge helpers
    "tiatoolbox/models/engine/multi_task_segmentor.py",  # text only: tile-merge functions
]


def _sha(p: Path) -> str:
    return hashlib.sha256(p.read_bytes()).hexdigest()[:16]


def vendor(quiet: bool = False) -> dict:
    """Copy FILES from /root/reference to oracle/_ref (no-op when the reference is absent, e.g. on the GPU box)."""
    if not (SRC / "tiatoolbox" / "tools" / "stainnorm.py").exists():
        return {}
    manifest = {}
    for rel in FILES:
        s, d = SRC / rel, DST / rel
        if not s.exists():
            continue
        d.parent.mkdir(parents=True, exist_ok=True)
        if not d.exists() or _sha(d) != _sha(s):
            shutil.copyfile(s, d)
        manifest[rel] = _sha(d)
    (DST / "MANIFEST.json").write_text(json.dumps(manifest, indent=1))
    if not quiet:
        print(f"[vendor_ref] {len(manifest)} reference files under {DST}")
    return manifest


def check() -> bool:
    ok = True
    for rel in FILES:
        s, d = SRC / rel, DST / rel
        if s.exists() and (not d.exists() or _sha(s) != _sha(d)):
            print("MISMATCH", rel)
            ok = False
    return ok


if __name__ == "__main__":
    if "--check" in sys.argv:
        sys.exit(0 if check() else 1)
    vendor()
