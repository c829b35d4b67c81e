#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT; T=r02s3
( timeout 300 python scripts/debug_polar_rb.py 2>&1 | tail -60 ) > $OUT/${T}_polar_rb.txt
( FH_POLAR_RB=0 timeout 300 python scripts/debug_polar_rb.py 12,137 2>&1 | tail -12 ) > $OUT/${T}_polar_rb0.txt
( FH_POLAR_RB=0 timeout 600 python bench.py --no-cpu-baseline --no-e2e 2>$OUT/${T}_bench_rb0.err | tail -1 ) > $OUT/${T}_bench_rb0.json
K='regex:avgpool_kernel|balance_kernel|chol_jacobi|col_norms_kernel|colsum_accum_kernel|delta_kernel|densify_conv_kernel|first_step_kernel|gemm_simt_kernel|gemm_tc_kernel|hadamard_inverse_kernel|identity_kernel|khatri_rao_kernel|mode0_reduce_kernel|mode1_reduce_kernel|ns_final_kernel|ns_init_kernel|ns_mid_kernel|rwr_chain_kernel|scale_cols_batched_kernel|sqnorm_kernel|symnorm_kernel|transition_kernel|triple_hadamard_sum_kernel'
FH_POLAR_RB=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 6000 --csv --log-file $OUT/${T}_launches.csv \
	python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 1 > $OUT/${T}_launches_bench.log 2>&1
python scripts/agg_launches.py $OUT/${T}_launches.csv $OUT/${T}_launches_bench.log > $OUT/${T}_launches_summary.txt 2>&1
gzip -f $OUT/${T}_launches.csv
cat $OUT/${T}_polar_rb.txt $OUT/${T}_polar_rb0.txt; head -40 $OUT/${T}_launches_summary.txt
python - <<'PY'
import json
try:
	d=json.load(open("gpurun_out/r02s3_bench_rb0.json")); print(d["ms_per_step"], d["stages_ms_per_sweep"])
except Exception as e: print("ERR", e)
PY
