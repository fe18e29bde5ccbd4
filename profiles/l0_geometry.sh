#!/bin/bash
# A/B sweep of the l0_kernel tile geometry (threads per CTA x resident CTAs per SM; alternative builds under build/, see
# PGR_L0_NT / PGR_L0_MIN_CTAS in shmmr_kernels.cuh) on 400 x 5 Mb contigs.  Run under gpurun; writes gpurun_out/.
out=gpurun_out/l0_geometry.jsonl
: > $out
for lib in build/libpgr_b200_nt256_c3.so build/libpgr_b200_nt128_c6.so build/libpgr_b200_nt384_c2.so; do
  for u in 4 8; do
    if [ $lib = default ]; then unset PGR_B200_LIB; else export PGR_B200_LIB=$PWD/$lib; fi
    export PGR_B200_L0_UNROLL=$u
    line=$(python bench.py --contigs 400 --steps 5 --warmup 3 --no-e2e 2>gpurun_out/err_geo.log | tail -1)
    echo "{\"lib\": \"$lib\", \"unroll\": $u, \"bench\": ${line:-null}}" >> $out
  done
done
python - <<'PY'
import json
for ln in open("gpurun_out/l0_geometry.jsonl"):
    d = json.loads(ln); b = d["bench"]
    if b: print(d["lib"], d["unroll"], "l0_ms %.3f" % b["roofline"]["kernel_ms"], "value %.1f" % b["value"], "cpu_parity", "cpu_baseline" in b)
    else: print(d["lib"], d["unroll"], "FAILED")
PY
for lib in build/libpgr_b200_nt384_c2.so; do
  PGR_B200_LIB=$PWD/$lib python -m pytest tests/test_gpu_shmmrs.py tests/test_gpu_properties.py -x -q -m gpu 2>&1 | tail -3
done

