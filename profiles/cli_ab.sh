#!/bin/bash
# wall clock of pgr-b200-make-frgdb on config 3 as 94 FASTA files (bench_cli.py writes them), plain against page-locked ingest slots
python bench_cli.py --config 3 --no-oracle --readers 8 > gpurun_out/cli_plain.jsonl 2> gpurun_out/cli_plain.err
CLI=pgr_tk_b200/pgr-b200-make-frgdb; FL=/tmp/pgr_b200_cli/files.txt
run() { local t0=$(date +%s%N); "$@" 2> /tmp/cli_err.txt; local t1=$(date +%s%N); echo "process_wall_ms $(( (t1 - t0) / 1000000 )) $(tail -1 /tmp/cli_err.txt)"; }
for r in 4 8; do for rep in 1 2 3; do echo "plain readers=$r: $(run $CLI $FL /tmp/pgr_b200_cli/o_plain --timing --readers $r --index-only)"; done; done
for r in 4 8; do for rep in 1 2 3; do echo "page-locked readers=$r: $(PGR_B200_INGEST_PINNED=1 run $CLI $FL /tmp/pgr_b200_cli/o_pin --timing --readers $r --index-only)"; done; done
cmp /tmp/pgr_b200_cli/o_plain.mdb /tmp/pgr_b200_cli/o_pin.mdb && echo MDB_IDENTICAL
cmp /tmp/pgr_b200_cli/o_plain.midx /tmp/pgr_b200_cli/o_pin.midx && echo MIDX_IDENTICAL
cmp /tmp/pgr_b200_cli/o_plain.mdb /tmp/pgr_b200_cli/out_g1_idx.mdb && echo MDB_IDENTICAL_TO_BENCH_RUN
echo "full (with the fragment store), plain, readers=8: $(run $CLI $FL /tmp/pgr_b200_cli/o_full --timing --readers 8)"
