"""The out-of-line exact-recompute calls of the suite kernels, checked in the SASS (tools/sass_call_check.py): no register that a
calling loop keeps live across the call may be in the write set of the callee chain.  (ptxas allocates registers across these
local calls itself; one experimental instantiation got it wrong -- DESIGN.md section 5.)  Needs the object files of the
in-tree build (git-ignored; present wherever __graft_entry__.build() ran) and cuobjdump."""
import glob
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("obj", ["ek_ops_fused_tqp.o", "ek_ops_fused_ttdp.o", "ek_ops_ept.o"])
def test_no_live_register_is_clobbered_across_the_cold_call(obj):
    import sass_call_check

    path = os.path.join(ROOT, "earthkit-meteo_b200", "csrc", "build", "lean", obj)
    if not os.path.exists(path) or shutil.which("cuobjdump") is None:
        pytest.skip("no object files of the in-tree build (or no cuobjdump) here")
    # the suites and the ept / wet-bulb kernels in float64: the kernels with the largest bodies (the fault showed in a suite)
    calls, bad = sass_call_check.check(path, r"ew_kernel.*EdLi2EEEvNS_6InArgs")
    assert calls > 50, calls
    assert bad == 0
