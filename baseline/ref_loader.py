"""Load the UNMODIFIED reference (installed offline into ``baseline/_ref`` with
``pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target
baseline/_ref /root/reference``) together with the ComfyUI stub.

``load()`` returns the reference module or raises ``ReferenceUnavailable`` with a
one-line reason (bench.py prints it as ``{"impl": "reference", "unavailable": ...}``).
"""
from __future__ import annotations

import importlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
STUB_DIR = os.path.join(HERE, "comfy_stub")
REF_SRC = "/root/reference"


class ReferenceUnavailable(RuntimeError):
    pass


def _try_install() -> None:
    if not os.path.isdir(REF_SRC):
        raise ReferenceUnavailable(f"{REF_DIR} missing and {REF_SRC} not present to install from")
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links",
           "/opt/wheelhouse", "--target", REF_DIR, REF_SRC]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    if r.returncode != 0:
        raise ReferenceUnavailable("pip install of the reference failed: " + r.stderr.strip().splitlines()[-1][:200])


def load():
    if not os.path.isfile(os.path.join(REF_DIR, "any_device_parallel.py")):
        _try_install()
    for p in (STUB_DIR, REF_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        import comfy.model_management  # noqa: F401  (real ComfyUI wins if it is importable)
        return importlib.import_module("any_device_parallel")
    except Exception as e:  # pragma: no cover
        raise ReferenceUnavailable(f"import of the reference failed: {type(e).__name__}: {e}")
