"""osep_path_planner_b200 -- B200 (sm_100a) implementation of osep_skeleton_decomp's per-frame
skeleton-extraction hot path behind the reference's own class interface (Rosa / GSkel).
The product is libosep_skel.so (csrc/, C ABI in include/osep_skel.h); this package is the thin
Python host binding used by the tests and bench.py.  The C++ shims with the reference's
member names are in osep_path_planner_b200/host/."""
from ._lib import RosaConfig, GSkelConfig, RosaResult, OsepError, load, LIB_PATH   # noqa: F401
from .rosa import Rosa   # noqa: F401
from .gskel import GSkel, build_graph   # noqa: F401
