"""oracle/ref_harness.py -- TEST INFRASTRUCTURE.

Imports the UNMODIFIED reference model from /root/reference (read-only, only present in the
build container) on top of the oracle shims for its two absent native dependencies
(``lib.pointnet2_utils`` and ``torch_scatter``).  Used by oracle/gen_golden.py to pin the
oracle port and produce tests/golden/.  Nothing run on the GPU box imports this module.
Recipe: SURVEY.md Appendix E.
"""
import os
import sys

REF_ASF = "/root/reference/scripts/ActiveSceneFlow"


def reference_available():
    return os.path.isfile(os.path.join(REF_ASF, "TFlowV3_Occlussion.py"))


def import_reference_tflow(module="TFlowV3_Occlussion"):
    """Returns the reference's ``TFlow`` class (unmodified source, executed on CPU) from the given model file."""
    if not reference_available():
        raise RuntimeError("reference tree not mounted at " + REF_ASF)
    sys.dont_write_bytecode = True  # the mount is read-only
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    for p in (REF_ASF, os.path.join(here, "shims"), root):
        if p not in sys.path:
            sys.path.insert(0, p)
    import importlib
    return importlib.import_module(module).TFlow
