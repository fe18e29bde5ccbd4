#!/bin/bash
# A/B of the level-0 kernel variants (run on the B200 box): device-resident config-2 step, 10 timed steps each, twice in
# alternation.  Variants are alternative builds of the library under profiles/ab/ (git-ignored; see the nvcc line in
# profiles/r2_l0_kernel_bulk_ab.txt), selected through PGR_B200_LIB.
for round in 1 2; do
for v in ${VARIANTS:-default plain}; do
  if [ $v = default ]; then unset PGR_B200_LIB; else export PGR_B200_LIB=$PWD/profiles/ab/libpgr_b200_$v.so; fi
  python - <<PY
import json, subprocess, sys, os
r = subprocess.run([sys.executable, "bench.py", "--no-index", "--no-assembly", "--no-e2e", "--no-sketch", "--steps", "10", "--warmup", "3"], capture_output=True, text=True)
try:
    d = json.loads(r.stdout.strip().splitlines()[-1])
    print("$v", "step_ms %.3f" % d["ms_per_step"], "l0_ms %.3f" % d["stages_ms"]["l0_minimizers"], "levels_ms %.3f" % d["stages_ms"]["reduce_and_span"], "Gbases/s %.1f" % d["value"], "parity(cpu sample)", "cpu_baseline" in d)
except Exception as e:
    print("$v FAILED", r.stderr[-400:])
PY
done
done
