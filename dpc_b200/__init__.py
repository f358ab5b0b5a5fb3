"""dpc_b200: B200-native (sm_100a) DPC-RNN training path behind the reference's nn.Module surface.

    from dpc_b200 import DPC_RNN, select_resnet, NCECriterion

Requires the in-tree CUDA library (python -m dpc_b200.build).  There is no CPU fallback.
"""
from ._lib import lib, DpcLibError, EXPORTS          # noqa: F401
from .select_backbone import select_resnet           # noqa: F401
from .model_3d import DPC_RNN                        # noqa: F401
from .criterion import NCECriterion                  # noqa: F401
from .parallel import FlatTrainer, shard_batch       # noqa: F401
from .model_3d_lc import LC                          # noqa: F401

__all__ = ['DPC_RNN', 'LC', 'select_resnet', 'NCECriterion', 'FlatTrainer', 'shard_batch', 'lib', 'DpcLibError']
