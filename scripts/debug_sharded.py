import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'oracle')); sys.path.insert(0,os.path.join(ROOT,'tests'))
import numpy as np
import oracle as O
from advancedps_b200 import _abi, _lib, models
from test_gpu_sharded import run_sharded
m = models.linear_gaussian(); N,T,world=4096,6,2
_, Y = O.simulate_data(m, T, 0xDA7A0005)
hs,out = run_sharded(m,N,T,Y,[11],world)
ro = O.sweep(_abi.make_config(m,N,T),Y,11,mode=O.CANON)
print("logev", out, ro.logevidence)
for r,h in enumerate(hs):
    logz,ess,rs=h.step_stats(); print(r,"logz",logz[:3],ro.logz[:3]); print(r,"ess",ess[:3],ro.ess[:3])
for t in range(1,T+1):
    x=np.concatenate([h.states(t) for h in hs]); bad=np.nonzero((x!=ro.x_hist[t-1]).any(axis=1))[0]
    print("t",t,"x bad",bad.size,bad[:6])
    if t>=2:
        a=np.concatenate([h.ancestors(t) for h in hs]); b=np.nonzero(a!=ro.anc_hist[t-1])[0]
        print("   anc bad",b.size,b[:6],a[b[:6]],ro.anc_hist[t-1][b[:6]])
    if bad.size: break
