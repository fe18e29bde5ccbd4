#!/bin/bash
# l0_kernel time against batch size (contigs of 5 Mb) for alternative builds: PGR_LIBS="a.so b.so" bash profiles/l0_sizes.sh "50 400"
for lib in ${PGR_LIBS:-default}; do
for n in ${1:-50 100 200 400 700 1000}; do
  if [ $lib = default ]; then unset PGR_B200_LIB; else export PGR_B200_LIB=$PWD/$lib; fi
  python bench.py --contigs $n --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.readline()); n=$n
print('$lib', n, 'l0_ms %.3f' % b['roofline']['kernel_ms'], 'ms/Gbase %.3f' % (b['roofline']['kernel_ms']/(n*5e6/1e9)), 'step_ms %.3f' % b['ms_per_step'])"
done
done
