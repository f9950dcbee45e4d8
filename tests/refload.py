"""Loads the UNMODIFIED reference extension built by oracle/build_ref.py (oracle/_ref/), if present."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def load_reference():
    """Returns (ransac_voting ext module, ransac_voting_gpu python module) or raises ImportError."""
    if not os.path.isdir(REF_DIR) or not any(f.endswith(".so") for f in os.listdir(REF_DIR)):
        raise ImportError("oracle/_ref not built (run python oracle/build_ref.py where /root/reference exists)")
    import torch  # noqa: F401  (the extension links against libtorch)
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    ext = importlib.import_module("ransac_voting")
    gpu = importlib.import_module("ransac_voting_gpu")
    return ext, gpu
