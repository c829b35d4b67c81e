#!/usr/bin/env bash
# One gpurun call that gathers everything a round needs from a B200 box, cheapest first, so that a
# cut-off call still leaves results behind in gpurun_out/:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_session.sh'
# 1. pytest -m gpu (verified tests), then the opt-in tests of code never run on a GPU (FH_RUN_UNVERIFIED=1)
# 2. smoke()
# 3. bench.py default line, reference arm, scripts/headline_run.py (the whole job) at N = 1
# 4. ncu launch list of a short bench (kernel shares) and one --set full capture of the fused RWR kernel
set -u
OUT=gpurun_out
mkdir -p $OUT
TAG=${1:-r02}
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $OUT/${TAG}_pytest_gpu.txt
( FH_RUN_UNVERIFIED=1 timeout 600 python -m pytest tests/test_zz_gpu_round2_candidates.py -m gpu -q 2>&1 | tail -40 ) > $OUT/${TAG}_pytest_unverified.txt
( timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 ) > $OUT/${TAG}_smoke.txt
( timeout 900 python bench.py 2>$OUT/${TAG}_bench_n1.err | tail -1 ) > $OUT/${TAG}_bench_n1.json
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 ) > $OUT/${TAG}_bench_reference_arm.json
# the whole job (init + sweeps + transform) of config 2 on this GPU, host and device init SVDs
( timeout 600 python scripts/headline_run.py --cells-total 4238 --sweeps 10 --geometry pfc --init-svd device 2>/dev/null | tail -1 ) > $OUT/${TAG}_job_n1_device_init.json
( timeout 600 python scripts/headline_run.py --cells-total 4238 --sweeps 10 --geometry pfc --init-svd host 2>/dev/null | tail -1 ) > $OUT/${TAG}_job_n1_host_init.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'densify|rwr_chain|gemm_|chol_jacobi|transition|khatri|mode|balance|hadamard|scale_cols|sqnorm|ns_' \
	-c 2200 --csv --log-file $OUT/${TAG}_launches_512cells.csv python bench.py --cells 512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/${TAG}_launches_bench.log 2>&1
python scripts/agg_launches.py $OUT/${TAG}_launches_512cells.csv > $OUT/${TAG}_launches_summary_512cells.txt 2>&1
gzip -f $OUT/${TAG}_launches_512cells.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rwr_chain_kernel -s 4 -c 1 -o $OUT/${TAG}_ncu_full_rwr_chain \
	python bench.py --cells 2072 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/${TAG}_ncu_full.log 2>&1
ncu -i $OUT/${TAG}_ncu_full_rwr_chain.ncu-rep --page details > $OUT/${TAG}_ncu_full_rwr_chain_kernel.txt 2>&1
# the opt-in block Jacobi of the per-bin polar step, timed next to the default line (only meaningful if its test above passed)
( FH_POLAR_BLOCK=1 timeout 600 python bench.py --no-cpu-baseline --no-e2e 2>$OUT/${TAG}_bench_polar_block.err | tail -1 ) > $OUT/${TAG}_bench_polar_block.json
# A/B of compile-time variants (nvcc is on the box): rebuild with the flag, parity test of the kernel, bench line; then restore
# (the chunk sizes change the summation order of the TF32 accumulations: the parity tests in front of the bench decide)
for V in FH_CHAIN_PROLOGUE_ROLLED FH_CHAIN_CHUNK_KB=5 FH_GEMM_CHUNK_KB=8; do
	( FH_NVCC_EXTRA="-D$V" timeout 600 python -c "import __graft_entry__ as g; g.build()" \
	  && timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "rwr or tcgen05 or lockstep or midsize" 2>&1 | tail -3 \
	  && FH_NVCC_EXTRA="-D$V" timeout 600 python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 ) > $OUT/${TAG}_variant_${V%%=*}.txt 2>&1
done
timeout 600 python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
tail -3 $OUT/${TAG}_pytest_gpu.txt $OUT/${TAG}_pytest_unverified.txt $OUT/${TAG}_smoke.txt
cat $OUT/${TAG}_bench_n1.json | cut -c1-600
