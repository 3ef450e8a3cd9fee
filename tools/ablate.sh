#!/bin/bash
# Timing-only ablations of the interpreter kernel on c3 (results of the runs with flags != 0 are wrong by design).
C=tests/golden/circuits/c3_surface_z_d25_r25.stim
export GSTIM_PHASED=${GSTIM_PHASED:-0}
for f in 0 2 1 3; do
  echo "== flags $f"
  GSTIM_DEBUG_FLAGS=$f python tools/prof_run.py $C 22 3 | tail -1
done
