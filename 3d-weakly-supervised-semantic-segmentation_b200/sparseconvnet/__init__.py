"""sparseconvnet -- B200-native drop-in for the `sparseconvnet` ("scn") operator layer used by
timsu1104/3D-Weakly-Supervised-Semantic-Segmentation (`import sparseconvnet as scn`,
models/SparseConvNet.py:5).  Same module names and call signatures; every sparse op runs as a
hand-written sm_100a CUDA kernel behind the C ABI in include/b200scn.h.  No CPU fallback, no Triton.
"""
import sys
import types

from . import _lib  # noqa: F401  (fails loudly if the CUDA library is missing)
from ._lib import B200SCNError, launch_count, set_option
from .metadata import Metadata, PackedKeys, set_pyramid_hint
from .modules import (AddTable, BatchNormalization, BatchNormLeakyReLU, BatchNormReLU, ConcatTable, Convolution,
                      Deconvolution, Identity, InputLayer, JoinTable, NetworkInNetwork, OutputLayer, Sequential,
                      SparseConvNetTensor, SubmanifoldConvolution, UnPooling)
from .networks import FullyConvolutionalNet, UNet
from .modules import MaxPooling, SceneMeanPooling, SparseToDense, set_fusion  # noqa: F401
from .ops import get_precision, invalidate_weight_cache, set_deferred_dw, set_precision, set_tiled, set_weight_cache
from .utils import checkpoint_restore, checkpoint_save, is_power2

forward_pass_hidden_states = 0


class _Module(types.ModuleType):
    """Module subclass so `scn.forward_pass_multiplyAdd_count` (train.py:50,86) stays a plain number for the
    caller while rule counts are resolved lazily: they travel device -> pinned host asynchronously and are folded in
    once their copy event has completed (never blocking a forward), or when somebody reads the counter.
    Only the small pinned count buffers are kept, never the rulebooks themselves."""

    def _fold(self, block):
        pend = self.__dict__["_madd_pending"]
        total = self.__dict__["_madd_base"]
        while pend:
            counts, event, mult = pend[0]
            if event is not None:
                if block:
                    event.synchronize()
                elif not event.query():
                    break
            total += int(counts.sum()) * mult if event is not None else int(counts) * mult
            pend.pop(0)
        self.__dict__["_madd_base"] = total
        return total

    @property
    def forward_pass_multiplyAdd_count(self):
        return self._fold(True)

    @forward_pass_multiplyAdd_count.setter
    def forward_pass_multiplyAdd_count(self, value):
        self.__dict__["_madd_base"] = value
        self.__dict__["_madd_pending"] = []

    def _add_madds(self, src, mult):
        pend = self.__dict__["_madd_pending"]
        if hasattr(src, "rule_counts"):
            src.subm_map()
            pend.append((src._counts_host, src._counts_event, mult))
        else:
            pend.append((int(src), None, mult))
        if len(pend) > 64:
            self._fold(False)


_self = sys.modules[__name__]
_self.__dict__["_madd_base"] = 0
_self.__dict__["_madd_pending"] = []
_self.__class__ = _Module
