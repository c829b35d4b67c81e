#!/usr/bin/env bash
# round-2 session 2: register-blocked polar kernel (tests, A/B timing, ncu), new parity tests, launch list at full size
set -u
OUT=gpurun_out; mkdir -p $OUT; T=r02s2
( timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=8 2>&1 | tail -60 ) > $OUT/${T}_pytest.txt
tail -4 $OUT/${T}_pytest.txt
( timeout 600 python bench.py --no-cpu-baseline --no-e2e 2>$OUT/${T}_bench_rb.err | tail -1 ) > $OUT/${T}_bench_rb.json
( FH_POLAR_RB=0 timeout 600 python bench.py --no-cpu-baseline --no-e2e 2>$OUT/${T}_bench_rb0.err | tail -1 ) > $OUT/${T}_bench_rb0.json
for f in $OUT/${T}_bench_rb.json $OUT/${T}_bench_rb0.json; do python - "$f" <<'PY'
import json,sys
try:
	d=json.load(open(sys.argv[1])); print(sys.argv[1], d["ms_per_step"], d["stages_ms_per_sweep"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
# kernel-level split of one sweep at full size (probe + 1 warm-up + 1 timed sweep)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/${T}_launches.csv \
	python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 1 > $OUT/${T}_launches_bench.log 2>&1
python scripts/agg_launches.py $OUT/${T}_launches.csv $OUT/${T}_launches_bench.log > $OUT/${T}_launches_summary.txt 2>&1; head -30 $OUT/${T}_launches_summary.txt
gzip -f $OUT/${T}_launches.csv
# full ncu of the largest polar class (first rb launch of a sweep = PL 5)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chol_jacobi_rb -s 5 -c 1 -o $OUT/${T}_ncu_polar_rb \
	python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 1 --cells 256 > $OUT/${T}_ncu_polar.log 2>&1
ls -la $OUT | tail -20
