#!/bin/bash
# C2 sweep (graph replay, device time) with the state draws made ahead of time on parallel graph branches
# (k_draw_normals) against the propagate kernel drawing them itself: steps ahead x fork point (1: behind the
# propagate kernel, 2: behind the normalise kernel) x blocks per SM x threads per block of the draw kernel.
# With a -DAPS_TIMELINE=1 build in $L and APS_DEBUG_MULTI=16 every line also carries the in-graph K1 / K2 / K3 spans.
L=${L:-advancedps.jl_b200/libaps_b200.so}
run() { python scripts/sweep_variants.py $L 2>&1 | grep -E "sweep ms|per step" | tail -2; }
echo "== no pre-draw";              APS_NO_PREDRAW=1 run
for a in ${AHEAD:-1 2}; do for f in ${FORK:-1 2}; do for b in ${BPS:-1}; do for th in ${THREADS:-128}; do
  echo "== pre-draw $a step(s) ahead, fork point $f, $b blocks per SM x $th threads"; APS_DRAW_AHEAD=$a APS_DRAW_FORK=$f APS_DRAW_BPS=$b APS_DRAW_THREADS=$th run; done; done; done; done
echo "== no pre-draw";              APS_NO_PREDRAW=1 run
