"""A/B of the C2 sweep (graph replay, device time) across library builds given on the command line."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    from advancedps_b200 import _abi, _lib, models
    import bench
    h = _lib.Handle(_abi.make_config(models.linear_gaussian(), int(os.environ.get("APS_SWEEP_N", 10**6)), 100))
    h.set_observations(bench.make_data())
    ms = []
    for k in range(13):
        le = h.sweep(1234 + k)
        if k >= 3:
            ms.append(h.last_sweep_ms())
    ms.sort()
    print(f"  sweep ms: min {ms[0]:.4f} median {ms[len(ms)//2]:.4f}  logev {le!r}")
else:
    for lib in sys.argv[1:]:
        print(lib, flush=True)
        subprocess.run([sys.executable, __file__, "--child"], env=dict(os.environ, APS_LIB_PATH=os.path.join(ROOT, lib)))
