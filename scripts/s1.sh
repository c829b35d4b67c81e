#!/usr/bin/env bash
# round-2 session 1: diagnostics of what round 1 left unverified on the GPU
set -u
OUT=gpurun_out; mkdir -p $OUT; T=r02s1
( FH_RUN_UNVERIFIED=1 timeout 900 python -m pytest tests/test_zz_gpu_round2_candidates.py -m gpu --runxfail -q --tb=short -p no:cacheprovider 2>&1 | tail -120 ) > $OUT/${T}_candidates.txt
( timeout 300 python scripts/debug_multires.py 2>&1 | tail -30 ) > $OUT/${T}_multires_default.txt
( timeout 300 python scripts/debug_multires.py --tc0 2>&1 | tail -30 ) > $OUT/${T}_multires_tc0.txt
( FH_RWR_FUSED=0 timeout 300 python scripts/debug_multires.py 2>&1 | tail -30 ) > $OUT/${T}_multires_fused0.txt
( FH_CP_STREAMS=0 timeout 300 python scripts/debug_multires.py 2>&1 | tail -30 ) > $OUT/${T}_multires_cpst0.txt
( timeout 600 python bench.py --no-cpu-baseline --no-e2e 2>$OUT/${T}_bench_default.err | tail -1 ) > $OUT/${T}_bench_default.json
( FH_POLAR_BLOCK=1 timeout 600 python bench.py --no-cpu-baseline --no-e2e 2>$OUT/${T}_bench_polar_block.err | tail -1 ) > $OUT/${T}_bench_polar_block.json
for V in FH_CHAIN_PROLOGUE_ROLLED; do
	( FH_NVCC_EXTRA="-D$V" timeout 600 python -c "import __graft_entry__ as g; g.build()" \
	  && FH_NVCC_EXTRA="-D$V" timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "rwr or lockstep" -p no:cacheprovider 2>&1 | tail -3 \
	  && FH_NVCC_EXTRA="-D$V" timeout 600 python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 ) > $OUT/${T}_variant_${V%%=*}.txt 2>&1
done
timeout 600 python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
tail -5 $OUT/${T}_candidates.txt
for f in $OUT/${T}_bench_default.json $OUT/${T}_bench_polar_block.json; do python - "$f" <<'PY'
import json,sys
try:
	d=json.load(open(sys.argv[1])); print(sys.argv[1], d["ms_per_step"], d["stages_ms_per_sweep"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
