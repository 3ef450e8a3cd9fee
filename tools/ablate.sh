#!/bin/bash
# Timing-only ablations of the interpreter kernel on c3 (results of the runs with flags != 0 are wrong by design).
C=tests/golden/circuits/c3_surface_z_d25_r25.stim
for f in 0 1 3072 16384 65536 130048; do
  echo "== flags $f"
  GSTIM_DEBUG_FLAGS=$f python tools/prof_run.py $C 22 3 | tail -1
done
echo "== producers 256"
GSTIM_PRE_THREADS=256 python tools/prof_run.py $C 22 3 | tail -1
echo "== producers 64"
GSTIM_PRE_THREADS=64 python tools/prof_run.py $C 22 3 | tail -1
