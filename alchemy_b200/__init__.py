"""alchemy_b200 -- B200-native backend for the ring arithmetic behind ALCHEMY's Lol `Tensor` class.

Python here is the host-side mirror of the reference's operator interface (Lol `class Tensor`,
SymmSHE's ciphertext operations as called from Crypto/Alchemy/Interpreter/Eval.hs); all arithmetic
runs in hand-written sm_100a CUDA kernels behind the C ABI of include/alchemy_b200.h.
"""
from ._lib import AlchemyError, INTERLEAVED, PLANAR, check, lib  # noqa: F401
from .tensor import GT, Ext, Plan  # noqa: F401
