"""clean_pvnet_b200 -- B200-native RANSAC voting layer (drop-in for clean-pvnet's lib/csrc/ransac_voting).

Public surface (same names/signatures as the reference):
    ransac_voting_gpu.ransac_voting_layer / ransac_voting_layer_v3 / estimate_voting_distribution_with_mean
    ransac_voting.generate_hypothesis / voting_for_hypothesis / *_vanishing_point   (the pybind twins)
    un_pnp.uncertainty_pnp / uncertainty_pnp_v2 (twins of lib/csrc/uncertainty_pnp/un_pnp_utils.py), uncertainty_pnp_batch,
    uncertainty_pnp_from_votes (the evaluator's whole un_pnp tail in one launch)
    parallel.ShardedVotingLayer (images sharded over the GPUs of one box, results exchanged over NVLink peer memory)
"""
from . import _lib  # noqa: F401
from . import ransac_voting  # noqa: F401
from . import ransac_voting_gpu  # noqa: F401
from . import decode  # noqa: F401
from . import parallel  # noqa: F401
from .decode import decode_keypoint, uncertainty_pnp_weights  # noqa: F401
from . import uncertainty_pnp as un_pnp  # noqa: F401
from .uncertainty_pnp import uncertainty_pnp_batch, p3p_init_batch, uncertainty_pnp_from_votes  # noqa: F401
from .ransac_voting_gpu import (  # noqa: F401
    estimate_voting_distribution_with_mean,
    ransac_voting_layer,
    ransac_voting_layer_v3,
    ransac_voting_layer_v3_host,
    install_as_reference_module,
)

__all__ = [
    "ransac_voting_layer", "ransac_voting_layer_v3", "estimate_voting_distribution_with_mean",
    "ransac_voting_layer_v3_host", "install_as_reference_module", "ransac_voting", "ransac_voting_gpu",
    "decode_keypoint", "uncertainty_pnp_weights", "un_pnp", "uncertainty_pnp_batch", "p3p_init_batch",
    "uncertainty_pnp_from_votes", "parallel",
]
