#!/bin/bash
# 2 GPUs: the time-sharded stream leg alone, dependent launch on / off / on (same box, interleaved)
mkdir -p gpurun_out
run() { timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-u8 --e2e-samples 20000000 --only-extra timeshard_chain 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); t = d['timeshard_chain']
print(json.dumps({'pdl': '$2', 'headline_ms': d['ms_per_step'], 'timeshard_ms': t['ms'], 'msps_total': t['msps_total'], 'exchange_halo_us': t['exchange_halo_us'], 'parity': t['timeshard_parity']}))"; }
run 29521 on | tee gpurun_out/r02_timeshard_pdl_ab.jsonl
DDM_CHAIN_NO_PDL=1 run 29522 off | tee -a gpurun_out/r02_timeshard_pdl_ab.jsonl
run 29523 on | tee -a gpurun_out/r02_timeshard_pdl_ab.jsonl
