"""maxk_b200 -- B200-native (sm_100a) MaxK-GNN sparse aggregation.

Host-side mirror of the reference's operator interface for the hot path
(TopK->CBSR, forward SpGEMM, backward SSpMM, work list) over the C ABI in
include/maxk_b200.h.  Importing this package does not load the CUDA library; the
first operator call does, and fails loudly if it is missing.
"""
from . import _cabi, ops  # noqa: F401
from .autograd import (CBSRAddScaled, CBSRAggregate, CBSRLinear, LinearTopKCBSR, MaxK, SpGEMMFunction,  # noqa: F401
                       TensorCoreLinear, TopKCBSR, worklist_for)

from .host import HostAggregator, ShardHostAggregator  # noqa: F401,E402

__all__ = ["ops", "MaxK", "SpGEMMFunction", "CBSRAggregate", "CBSRLinear", "CBSRAddScaled", "LinearTopKCBSR", "TensorCoreLinear", "TopKCBSR", "worklist_for",
           "HostAggregator", "ShardHostAggregator"]
