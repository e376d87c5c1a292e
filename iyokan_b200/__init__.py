"""iyokan_b200 — B200-native TFHE gate-evaluation back-end for Iyokan (hot path only).

The product is the CUDA library behind ``include/b200fhe.h`` (``csrc/``) plus the host-side
mirror of Iyokan's back-end interface.  See DESIGN.md.
"""
from .lib import B200FheError, Context, PinnedBuffer, OPS, BOOTSTRAPS, load  # noqa: F401

__all__ = ["B200FheError", "Context", "PinnedBuffer", "OPS", "BOOTSTRAPS", "load"]
