"""Times the isolated resample kernel for each library variant given on the command line."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    from advancedps_b200 import _abi, _lib
    for n, fl in ((1 << 25, 2), (1 << 25, 1), (1 << 25, 0), (1 << 20, 2)):
        avg, mn = _lib.bench_resample(_abi.RESAMPLE_SYSTEMATIC, n, iters=20, flush_l2=fl)
        print(f"  n={n} flush={fl}: avg {avg*1e3:.1f} us  min {mn*1e3:.1f} us  -> {12*n/(avg*1e-3)/1e9:.0f} GB/s alg")
else:
    for lib in sys.argv[1:]:
        print(lib, flush=True)
        env = dict(os.environ, APS_LIB_PATH=os.path.join(ROOT, lib))
        subprocess.run([sys.executable, __file__, "--child"], env=env)
