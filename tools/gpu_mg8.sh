#!/bin/bash
# 8 GPUs: the bench at N = 8 and N = 4 (strong scaling of the 8M workload), then the 32M pacing workload at N = 8
mkdir -p gpurun_out
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_mg_$N.log 2>gpurun_out/bench_mg_$N.err; echo "bench $N exit $?"
grep "^{" gpurun_out/bench_mg_$N.log | python -c "
import sys,json
for ln in sys.stdin:
    d=json.loads(ln); print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e']['h2d_bytes_per_step'], d['e2e']['d2h_bytes_per_step'], d.get('mg_parity'), d['roofline']['kernel_group_ms'])
"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus 8 --steps 100 --warmup 5 --workload 32m --no-cpu-baseline > gpurun_out/bench_mg_8_32m.log 2>&1; echo "bench 32m exit $?"
grep "^{" gpurun_out/bench_mg_8_32m.log | python -c "
import sys,json
for ln in sys.stdin:
    d=json.loads(ln); print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, 'e2e', d['e2e']['value'], d['roofline']['whole_step'])
"
