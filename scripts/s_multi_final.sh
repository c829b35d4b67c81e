#!/usr/bin/env bash
# end-of-round-2 multi-GPU session (one 8-GPU box): N-GPU == 1-GPU parity record and the headline job (config 4) with the
# 3xFP16 RWR kernel
set -u
OUT=gpurun_out; mkdir -p $OUT; T=${1:-r02fm}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( NCCL_DEBUG=INFO timeout 300 $TR --nproc-per-node 8 --master-port 29501 scripts/multigpu_check.py 2>&1 | grep -E "rank |MULTIGPU|NVLS|Error|error|Traceback" | head -60 ) > $OUT/${T}_multigpu_check_n8.txt
tail -3 $OUT/${T}_multigpu_check_n8.txt
( timeout 600 $TR --nproc-per-node 8 --master-port 29502 scripts/headline_run.py --cells-total 100000 --sweeps 60 2>$OUT/${T}_headline.err | tail -1 ) > $OUT/${T}_headline_config4_n8.json
cut -c1-1200 $OUT/${T}_headline_config4_n8.json; tail -2 $OUT/${T}_headline.err
