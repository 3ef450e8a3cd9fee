#!/bin/bash
# Launch-parameter sweep of the interpreter on c3 (and the other configs at their defaults).
C=tests/golden/circuits
run() { echo "== $*"; env "$@" python tools/prof_run.py $C/c3_surface_z_d25_r25.stim 22 3 | tail -1; }
run GSTIM_PRE_THREADS=128
run GSTIM_PRE_THREADS=96
run GSTIM_PRE_THREADS=160
run GSTIM_SLOTS=608
run GSTIM_SLOTS=672 GSTIM_PRE_THREADS=96
run GSTIM_CHUNK_WORDS=2048
run GSTIM_CHUNK_WORDS=2560
for c in c1_rep_d3_r10 c2_surface_x_d5_r5 c4_color_d15_r15 c4v_color_d15_r15_mpp_dense; do echo "== $c"; python tools/prof_run.py $C/$c.stim 22 3 | tail -1; done
echo "== c5"; python tools/prof_run.py $C/c5_surface_x_d51_r51.stim 20 3 | tail -1
echo "== c5 phased"; GSTIM_PHASED=1 python tools/prof_run.py $C/c5_surface_x_d51_r51.stim 20 3 | tail -1
