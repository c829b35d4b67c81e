#!/usr/bin/env bash
# end-of-round-2 multi-GPU session, part 2 (one 8-GPU box, final kernels): weak-scaling bench line at N = 8 and the
# strong-scaling curve of config 3 (N = 8, then 4 / 2 / 1 side by side on disjoint GPUs)
set -u
OUT=gpurun_out; mkdir -p $OUT; T=${1:-r02fm}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( timeout 600 $TR --nproc-per-node 8 --master-port 29503 bench.py --gpus 8 --no-cpu-baseline --torch-gpu-sample-cells 0 2>$OUT/${T}_bench_n8.err | tail -1 ) > $OUT/${T}_bench_n8.json
cut -c1-400 $OUT/${T}_bench_n8.json; tail -2 $OUT/${T}_bench_n8.err
( timeout 600 $TR --nproc-per-node 8 --master-port 29504 scripts/headline_run.py --cells-total 20000 --sweeps 12 2>$OUT/${T}_cfg3_n8.err | tail -1 ) > $OUT/${T}_config3_n8.json
( CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 600 $TR --nproc-per-node 4 --master-port 29505 scripts/headline_run.py --cells-total 20000 --sweeps 12 2>$OUT/${T}_cfg3_n4.err | tail -1 ) > $OUT/${T}_config3_n4.json &
( CUDA_VISIBLE_DEVICES=4,5 timeout 600 $TR --nproc-per-node 2 --master-port 29506 scripts/headline_run.py --cells-total 20000 --sweeps 12 2>$OUT/${T}_cfg3_n2.err | tail -1 ) > $OUT/${T}_config3_n2.json &
( CUDA_VISIBLE_DEVICES=6 timeout 600 python scripts/headline_run.py --cells-total 20000 --sweeps 12 2>$OUT/${T}_cfg3_n1.err | tail -1 ) > $OUT/${T}_config3_n1.json &
wait
for n in 8 4 2 1; do python - "$OUT/${T}_config3_n$n.json" <<'PY'
import json,sys
try:
	d=json.load(open(sys.argv[1])); print(d["n_gpus"], "init", round(d["init_s"],2), "sweep", round(d["sweep_median_s"]*1e3,1), "ms", int(d["cells_per_s_per_sweep"]), "cells/s", d["stages_ms_per_sweep_last"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
