"""oracle/shims/torch_scatter -- TEST INFRASTRUCTURE: stand-in for the unvendored torch_scatter
wheel (call sites ASF/utils/soflow.py:13,474,481; 3-D src, dim=1 only)."""
from oracle.point_ops import scatter_softmax, scatter_sum  # noqa: F401
