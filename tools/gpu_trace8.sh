#!/bin/bash
# 8 GPUs: event timeline of three slab steps in the middle of the timed region (SPHSM_TRACE), then the plain bench line
mkdir -p gpurun_out
SPHSM_TRACE=450 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/trace8.log 2> gpurun_out/trace8.err; echo "exit $?"
grep -A70 "sphsm trace rank 0" gpurun_out/trace8.err | head -75
grep -A70 "sphsm trace rank 4" gpurun_out/trace8.err | head -75
grep "^{" gpurun_out/trace8.log | python -c "
import sys,json
for ln in sys.stdin:
    d=json.loads(ln); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], d['mg_parity']['bit_identical'])
"
