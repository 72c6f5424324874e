"""``lib.pointnet2_utils`` for unmodified reference files: re-exports the B200 kernels' Python surface."""
from ssf_slam_b200.pointnet2_utils import *  # noqa: F401,F403
from ssf_slam_b200.pointnet2_utils import (GroupAll, QueryAndGroup, ball_query, furthest_point_sample,  # noqa: F401
                                           gather_operation, grouping_operation, knn, three_interpolate, three_nn)
