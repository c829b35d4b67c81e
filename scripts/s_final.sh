#!/usr/bin/env bash
# end-of-round-2 measurement session on one B200: launch list of one timed sweep, ncu --set full of the fused RWR kernel and
# of the binary16 densify, the default bench line (CPU and stock-torch baselines included)
set -u
OUT=gpurun_out; mkdir -p $OUT; T=${1:-r02f}
K='regex:avgpool_kernel|balance_kernel|chol_jacobi|col_norms_kernel|colsum_accum_kernel|cp_commit_kernel|cp_stop_kernel|csr_absmax_kernel|delta_kernel|densify_conv_kernel|first_step_kernel|gemm_simt_kernel|gemm_tc_kernel|hadamard_inverse_kernel|identity_kernel|khatri_rao_kernel|mode0_reduce_kernel|mode1_reduce_kernel|ns_fused_kernel|rwr_chain|scale_cols_batched_kernel|sqnorm_kernel|symnorm_kernel|transition_kernel|triple_hadamard_sum_kernel'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 9000 --csv --log-file $OUT/${T}_launches.csv \
	python bench.py --no-cpu-baseline --no-e2e --torch-gpu-sample-cells 0 --steps 1 --warmup 1 > $OUT/${T}_launches_bench.log 2>&1
python scripts/agg_launches.py $OUT/${T}_launches.csv $OUT/${T}_launches_bench.log > $OUT/${T}_launches_summary.txt 2>&1
gzip -f $OUT/${T}_launches.csv
head -24 $OUT/${T}_launches_summary.txt
# the first fused-kernel launches of a sweep are chr1's first blocks (nb 115, w 215 / 315); bench.py's RWR-step probe launches
# come first (99 of each kernel)
for spec in "rwr_chain16_kernel:100:chain16" "densify_conv_kernel:100:densify16"; do
	IFS=: read kn skip tag <<< "$spec"
	timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kn -s $skip -c 1 -f -o $OUT/${T}_ncu_$tag \
		python bench.py --no-cpu-baseline --no-e2e --torch-gpu-sample-cells 0 --steps 1 --warmup 1 > $OUT/${T}_ncu_$tag.log 2>&1
	ncu -i $OUT/${T}_ncu_$tag.ncu-rep --page details > $OUT/${T}_ncu_$tag.txt 2>&1
	grep -E "^  [a-z_<>0-9:, ]+\(|Duration|DRAM Throughput|Memory Throughput|Compute \(SM\)" $OUT/${T}_ncu_$tag.txt | head -8
done
timeout 900 python bench.py > $OUT/${T}_bench.json 2> $OUT/${T}_bench.err
tail -c 600 $OUT/${T}_bench.json
ls -la $OUT | grep $T
