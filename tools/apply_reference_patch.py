#!/usr/bin/env python
"""Copies a checkout of the reference package and applies integration/roger_cuda_backend.patch to the COPY:

    python tools/apply_reference_patch.py /path/to/reference /tmp/roger_cuda      # -> /tmp/roger_cuda/roger (patched)

Put the target directory in front of PYTHONPATH (together with this repository) and select the backend with
ROGER_BACKEND=cuda, `roger-run -b cuda` or `runtime_settings.update(backend="cuda")`."""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATCH = os.path.join(ROOT, "integration", "roger_cuda_backend.patch")


def apply(reference_root, target):
    os.makedirs(target, exist_ok=True)
    dst = os.path.join(target, "roger")
    if os.path.exists(dst):
        shutil.rmtree(dst)
    shutil.copytree(os.path.join(reference_root, "roger"), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    subprocess.run(["patch", "-p1", "--quiet", "-i", PATCH], cwd=target, check=True)
    return dst


if __name__ == "__main__":
    print(apply(sys.argv[1], sys.argv[2]))
