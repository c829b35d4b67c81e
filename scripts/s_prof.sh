#!/usr/bin/env bash
# round-2 profiling session: launch list of one timed sweep at full size + ncu --set full of the four hot kernels
set -u
OUT=gpurun_out; mkdir -p $OUT; T=r02p1
K='regex:avgpool_kernel|balance_kernel|chol_jacobi|col_norms_kernel|colsum_accum_kernel|cp_commit_kernel|cp_stop_kernel|delta_kernel|densify_conv_kernel|first_step_kernel|gemm_simt_kernel|gemm_tc_kernel|hadamard_inverse_kernel|identity_kernel|khatri_rao_kernel|mode0_reduce_kernel|mode1_reduce_kernel|ns_fused_kernel|rwr_chain_kernel|scale_cols_batched_kernel|sqnorm_kernel|symnorm_kernel|transition_kernel|triple_hadamard_sum_kernel'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 8000 --csv --log-file $OUT/${T}_launches.csv \
	python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 1 > $OUT/${T}_launches_bench.log 2>&1
python scripts/agg_launches.py $OUT/${T}_launches.csv $OUT/${T}_launches_bench.log > $OUT/${T}_launches_summary.txt 2>&1
gzip -f $OUT/${T}_launches.csv
head -32 $OUT/${T}_launches_summary.txt
# full captures: skip the RWR-step probe (99 densify / chain launches of bench.py's probe_rwr_steps use the per-op path) so
# that the captured launch belongs to a sweep; the first launches of a sweep are chr1's first block (nb 115, w 215 / 315)
for spec in "rwr_chain_kernel:1:chain" "densify_conv_kernel:55:densify" "chol_jacobi_rb:0:jacobi" "gemm_tc_kernel:164:gemm"; do
	IFS=: read kn skip tag <<< "$spec"
	timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kn -s $skip -c 1 -o $OUT/${T}_ncu_$tag \
		python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 1 > $OUT/${T}_ncu_$tag.log 2>&1
	ncu -i $OUT/${T}_ncu_$tag.ncu-rep --page details > $OUT/${T}_ncu_$tag.txt 2>&1
	grep -E "^  [a-z_<>0-9:, ]+\(|Duration|DRAM Throughput|Memory Throughput|Compute \(SM\)" $OUT/${T}_ncu_$tag.txt | head -8
done
ls -la $OUT | grep $T
