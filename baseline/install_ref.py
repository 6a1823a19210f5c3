"""TEST / BASELINE INFRASTRUCTURE: "installs" the unmodified reference for the GPU box.

The reference (tum-vision/scenedino) has no setup.py / pyproject.toml, so ``pip install --target baseline/_ref /root/reference``
has nothing to build; its Python tree (``*.py`` and the ``*.yaml`` files its modules open at import) is copied as is into
``baseline/_ref`` instead.  That directory is git-ignored (no reference source enters the repository's history) but not
gpurun-ignored, so it travels to the GPU box with the snapshot -- where /root/reference does not exist -- for
  * the drop-in tests (tests/test_gpu_dropin.py: the reference's own ``inference_rendered_2d`` / ``inference_3d`` /
    ``downsample_and_predict`` on top of the B200-native classes), and
  * ``bench.py --impl reference`` / ``cpu_baseline`` with ``kind: "reference"`` (the reference's PyTorch CPU path).
Run by ``__graft_entry__.build()`` when /root/reference is present; a no-op otherwise.

    python baseline/install_ref.py
"""
from __future__ import annotations

import os
import shutil
import sys

SRC = os.environ.get("SCENEDINO_REFERENCE", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
KEEP = (".py", ".yaml", ".yml")


def install(force: bool = False) -> str | None:
    if not os.path.isdir(os.path.join(SRC, "scenedino")):
        return DST if os.path.isdir(os.path.join(DST, "scenedino")) else None
    marker = os.path.join(DST, ".installed")
    if os.path.exists(marker) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    n = 0
    for root, dirs, files in os.walk(SRC):
        dirs[:] = [d for d in dirs if not d.startswith(".") and d != "__pycache__"]
        rel = os.path.relpath(root, SRC)
        for f in files:
            if f.endswith(KEEP):
                os.makedirs(os.path.join(DST, rel), exist_ok=True)
                shutil.copyfile(os.path.join(root, f), os.path.join(DST, rel, f))
                n += 1
    with open(marker, "w") as fh:
        fh.write(f"{n} files copied from {SRC}\n")
    return DST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
