"""Test stand-in for pyrootutils: the real setup_root puts the project root of `search_from` at the
FRONT of sys.path and exports PROJECT_ROOT; so does this one."""
import os
import sys
from pathlib import Path


def setup_root(search_from, indicator=".project-root", pythonpath=True, **kwargs):
    p = Path(search_from).resolve()
    root = next((d for d in [p, *p.parents] if (d / indicator).exists()), p.parent)
    os.environ["PROJECT_ROOT"] = str(root)
    if pythonpath:
        sys.path.insert(0, str(root))
    return root
