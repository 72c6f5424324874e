"""``lib`` package as the reference imports it (``from lib import pointnet2_utils``, ASF/utils/utils.py:7)."""
