"""Drop-in for the reference's ``cross_atten/pscan.py`` (same names, same argument meaning).

``pscan(A, X)`` runs the sm_100a streaming-scan kernels (csrc/pscan.cu) instead of the Blelloch op stream of
pscan.py:37-149.  ``npo2`` / ``pad_npo2`` are kept because callers may import them (pscan.py:13-33); the
kernels themselves need no padding -- the reference pads after position L-1, which never changes [0, L).
"""
import math

import torch
import torch.nn.functional as F

from .ops import _PScanFn


def npo2(len):
    """Next power of two >= len (pscan.py:13-18)."""
    return 2 ** math.ceil(math.log2(len))


def pad_npo2(X):
    """Zero-pad dim 1 of a (B, L, D, N) tensor to the next power of two (pscan.py:20-33)."""
    len_npo2 = npo2(X.size(1))
    return F.pad(X, (0, 0, 0, 0, 0, len_npo2 - X.size(1)), "constant", 0)


class PScan(_PScanFn):
    """Same role as the reference's ``PScan`` autograd Function (pscan.py:35): ``PScan.apply(A, X) -> H`` with
    A, X, H of shape (B, L, D, N); backward returns (dA, dX) per pscan.py:189-224."""


pscan = PScan.apply
