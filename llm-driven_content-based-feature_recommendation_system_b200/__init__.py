"""B200-native (sm_100a) implementation of the two-tower training / retrieval hot path of
DotBlossom/LLM-driven_content-based-feature_recommendation_system.

Import loads librs_twotower.so (C ABI: include/rs_twotower.h) and registers the `torch.ops.rs.*`
custom ops.  There is no CPU or eager-PyTorch fallback: a missing library raises here.
"""
from . import _lib
from ._lib import check_ids, LIB_PATH

_lib.load()

from . import ops                                               # noqa: E402  (registers torch.ops.rs.*)
from .ops import (gather_rows, seq_front, static_front, normalized_rows, masked_mean, fm_interaction)  # noqa: E402
from . import encoder, losses, towers, fm, retrieval, train, sharded, synthetic     # noqa: E402
from . import alignment, ensemble, lightgcl                                        # noqa: E402
from .losses import (simcse_loss, inbatch_corrected_logq_loss, inbatch_logq_loss_no_user, duorec_loss_refined,  # noqa
                     logq_correction_loss, efficient_corrected_logq_loss, logq_infonce_rows, info_nce,
                     logq_infonce_columns, item_columns,
                     full_batch_hard_emphasis_loss, inbatch_hnm_corrected_loss_with_stats,
                     inbatch_mixed_hnm_loss_with_stats)
from .towers import (SASRecUserTower, SASRecItemTower, HybridItemTower, OptimizedItemTower, SimCSEModelWrapper,  # noqa
                     HybridUserEmbeddings, DeepResidualHead)
from .fm import FM, DeepFM                                      # noqa: E402
from .retrieval import retrieve_topk                            # noqa: E402

__all__ = [n for n in dir() if not n.startswith("_")]
