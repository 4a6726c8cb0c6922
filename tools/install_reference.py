#!/usr/bin/env python
"""
Install the UNMODIFIED reference (vllm-project/compressed-tensors, /root/reference) into baseline/_ref/ so that bench.py's
reference arm can time the reference's own code on the GPU box's host cores (baseline/_ref is git-ignored, not gpurun-ignored: it
travels with the snapshot; /root/reference itself does not exist on the GPU box).

    python tools/install_reference.py [--force]

`pip install --no-index --no-build-isolation --target baseline/_ref /root/reference` fails in this image (setup.py imports
setuptools_scm, which is not installed and there is no network), so this is BASELINE.md section 4's recipe: copy the package directory
as it lies under /root/reference/src and write the two-line version.py that setup.py would have generated
(src/compressed_tensors/__init__.py:22 imports it).  Nothing else is touched; no reference source enters the git history.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/src/compressed_tensors"
DST = os.path.join(ROOT, "baseline", "_ref", "compressed_tensors")


def install(force: bool = False) -> str:
    """returns 'installed' | 'present' | 'unavailable: <why>'"""
    if os.path.exists(os.path.join(DST, "version.py")) and not force:
        return "present"
    if not os.path.isdir(SRC):
        return "unavailable: /root/reference is not on this machine"
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__"))
    with open(os.path.join(DST, "version.py"), "w") as f:
        f.write('__version__ = version = "0.0.0+ref"\n__all__ = ["__version__", "version"]\n')
    return "installed"


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
