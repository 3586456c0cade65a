"""ev2hands_b200 - the Ev2Hands set-abstraction encoder hot path as sm_100a CUDA
kernels behind a C ABI (include/ev2h.h), with drop-in PyTorch modules."""
from .pointnet2_utils import (  # noqa: F401
    PointNetFeaturePropagation,
    PointNetSetAbstraction,
    PointNetSetAbstractionMsg,
    farthest_point_sample,
    get_mlp_precision,
    set_mlp_precision,
    index_points,
    query_ball_point,
    sample_and_group,
    sample_and_group_all,
    square_distance,
)
from ._capi import check_numeric_range  # noqa: F401
from .windows import EventWindowBuilder  # noqa: F401
from .encoder import FeaturePropagationDecoder, RegressorSetAbstraction, SetAbstractionEncoder  # noqa: F401

__version__ = "0.1.0"
