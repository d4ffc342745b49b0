"""b200ot -- B200-native entropic optimal transport for the MRI<->PET alignment hot path.

Host side of the C ABI in include/b200ot.h.  ``b200ot.api`` carries the reference's call
surface (``sinkhorn``, ``sinkhorn_scaling``, ``linear_solve``/``Geometry``,
``get_feature_coupling_pot``, ``get_coupling_fot``, ...); ``b200ot.ops`` the device-level
operators; ``b200ot.torch_ops`` the ``torch.library`` custom ops and the autograd function;
``b200ot.sharded`` the row-sharded multi-GPU solver.  There is no CPU fallback.
"""
from ._lib import B200OTError, LIB_PATH  # noqa: F401
from . import ops  # noqa: F401
from .api import (Geometry, SinkhornOutput, dist, unif, foscttm, get_FOSCTTM, group_features_by_label, fot_numpy, get_coupling_fot, get_feature_coupling_pot,  # noqa: F401
                  get_coupling_egw_ott_fixed, get_coupling_egw_ott, get_coupling_eot_ott, get_coupling_egw_all_ott,
                  get_coupling_egw_labels_ott, get_coupling_leot_ott, per_step_feature_plan, cotl_numpy, get_coupling_cotl_sinkhorn, compute_pet_to_mri_coupling,
                  init_matrix_np, linear_solve, mdict_to_matrix, sinkhorn, sinkhorn_from_embeddings,
                  sinkhorn_scaling)

__version__ = "0.1.0"
