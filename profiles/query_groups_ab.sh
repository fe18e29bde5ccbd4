#!/bin/bash
# config 4 (10 000 x 20 kb queries against the 94-haplotype map) against the number of query groups of pgr_b200_query_batch
# (PGR_B200_QUERY_GROUPS; default = 4 for a batch of this size): ms per call and the 1000-query parity flag of bench_configs.py
for g in default 1 2 4 8; do
  if [ $g = default ]; then unset PGR_B200_QUERY_GROUPS; else export PGR_B200_QUERY_GROUPS=$g; fi
  python bench_configs.py --configs 3,4 2>/dev/null | grep '"config": 4' | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('groups $g: %.2f ms  times %s  hit pairs %d  chains %d  parity %s' % (d['ms'], [round(x,1) for x in d['times_ms']], d['hit_pairs'], d['chains'], d['parity_1000_queries_vs_oracle_same_index']))"
done
