"""Golden vectors produced by running the reference (see make_golden.py)."""
import gzip
import io
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def load(name: str):
    """load tests/golden/<name>.pt.gz (plain tensors / python scalars only)"""
    with gzip.open(os.path.join(_HERE, name + ".pt.gz"), "rb") as f:
        return torch.load(io.BytesIO(f.read()), weights_only=True)
