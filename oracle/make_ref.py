"""Test / measurement infrastructure, not product code: puts an UNMODIFIED copy of the reference's Python sources
(spatial_vae/, src/, train_*.py of cfframe/spatial-VAE) under baseline/_ref/ so that bench.py's reference arm and its
cpu_baseline / gpu_eager_baseline legs can time the reference itself on the GPU box, where /root/reference does not exist.

baseline/_ref/ is git-ignored (never part of the history) and NOT gpurun-ignored (it travels with the snapshot).  The
reference has no setup.py / pyproject, so `pip install --target baseline/_ref /root/reference` has nothing to build:
this copy is the equivalent.  Called by __graft_entry__.build() when /root/reference is present.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SVAE_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def make_ref() -> bool:
    if not os.path.isdir(os.path.join(REF, "spatial_vae")):
        return os.path.isdir(os.path.join(DST, "spatial_vae"))
    os.makedirs(DST, exist_ok=True)
    for name in ("spatial_vae", "src"):
        shutil.copytree(os.path.join(REF, name), os.path.join(DST, name), dirs_exist_ok=True,
                        ignore=shutil.ignore_patterns("__pycache__"))
    for name in ("train_mnist.py", "train_particles.py", "train_galaxy.py", "LICENSE"):
        if os.path.exists(os.path.join(REF, name)):
            shutil.copy2(os.path.join(REF, name), os.path.join(DST, name))
    return True


if __name__ == "__main__":
    print("baseline/_ref ready" if make_ref() else "reference not available", file=sys.stderr)
