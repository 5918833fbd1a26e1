"""gspn_b200 -- B200-native PointNet++ set-abstraction / feature-propagation engine behind the
op surface of ericyi/GSPN (tf_ops + utils/pointnet_util.py).  See DESIGN.md."""
from . import _lib  # noqa: F401
from .ops import (farthest_point_sample, gather_point, query_ball_point, group_point, three_nn, three_interpolate,  # noqa: F401
                  nn_distance, nearest_point, nearest_point_index, box_shrink)
from .pointnet_util import pointnet_sa_module, pointnet_fp_module, sample_and_group  # noqa: F401

__all__ = ["farthest_point_sample", "gather_point", "query_ball_point", "group_point", "three_nn", "three_interpolate",
           "nn_distance", "nearest_point", "nearest_point_index", "box_shrink", "pointnet_sa_module", "pointnet_fp_module", "sample_and_group"]
