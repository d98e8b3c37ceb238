"""Small sweeps for compute-sanitizer (memcheck) of the round-2b kernels: k_draw_normals on its parallel stream and
k_propagate<..., PRE> / k_propagate1<..., PRE> (APS_PREDRAW=1 forces the path at these sizes), one and two steps ahead."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
os.environ["APS_PREDRAW"] = "1"; os.environ["APS_NO_FUSED"] = "1"
import oracle as O
from advancedps_b200 import _abi, _lib, models

def run(tag, m, N, T, ahead, **kw):
    os.environ["APS_DRAW_AHEAD"] = str(ahead)
    _, Y = O.simulate_data(m, T, 7)
    cfg = _abi.make_config(m, N, T, **kw)
    h = _lib.Handle(cfg); h.set_observations(Y)
    ro = O.sweep(cfg, Y, 3, mode=O.CANON)
    le = h.sweep(3)
    ok = le == ro.logevidence and np.array_equal(h.ancestors(T + 1), ro.anc_hist[T])
    print(tag, "ahead", ahead, "launches", h.last_sweep_launches(), "OK" if ok else "MISMATCH", flush=True)

lg = models.linear_gaussian()
run("lg1 sys N=20011", lg, 20011, 4, 1)
run("lg1 sys N=20011", lg, 20011, 4, 2)
run("lg1 ess N=4097", lg, 4097, 3, 1, ess_threshold=0.5)
run("lg4 N=5001", models.lg4(), 5001, 3, 1)
