"""``torch_scatter`` stand-in for unmodified reference files (ASF/utils/soflow.py:13)."""
from ssf_slam_b200.scatter import scatter_softmax, scatter_sum  # noqa: F401
