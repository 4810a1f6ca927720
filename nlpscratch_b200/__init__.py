"""nlpscratch_b200 -- the B200 (sm_100a) training step of the NLPScratch Transformer.

    from nlpscratch_b200 import Transformer, b200np

``Transformer`` mirrors ``Transformer/transformer.py::Transformer`` of the reference; ``b200np`` is
the array module it exposes as ``model.np``.  All arithmetic runs in hand-written CUDA kernels
behind ``libnlps_b200.so`` (``include/nlps_b200.h``); importing the package never touches torch.
"""
from . import _lib
from .array import DeviceArray, b200np, matmul
from .transformer import PARAM_NAMES, Cache, DeviceScalar, HostScalarSlot, Transformer

__all__ = ["Transformer", "DeviceArray", "DeviceScalar", "HostScalarSlot", "Cache", "b200np", "matmul", "PARAM_NAMES", "_lib"]
__version__ = "0.1.0"
