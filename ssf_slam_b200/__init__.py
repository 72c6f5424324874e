"""ssf_slam_b200 -- B200-native (sm_100a) scene-flow front end of SSF-SLAM.

Public surface (mirrors the reference's Python call surface for this path):
  ssf_slam_b200.pointnet2_utils   drop-in for `lib.pointnet2_utils`      (B-op)
  ssf_slam_b200.scatter           drop-in for `torch_scatter`            (B-scatter)
  ssf_slam_b200.model.TFlow       drop-in for `TFlowV3_Occlussion.TFlow` (B-model / B-layer)
  ssf_slam_b200.frontend          slove_RT_by_SVD, background_index, odometry, SceneFlowFrontEnd (B-frontend)
  ssf_slam_b200.wire              PointCloud2 / Float64MultiArray byte layouts (B-wire)
  ssf_slam_b200.compat/           `lib` and `torch_scatter` packages for unmodified reference files
All compute runs in hand-written CUDA kernels behind the C ABI in include/ssf_b200.h; nothing falls back to CPU.
"""
__version__ = "0.1.0"
