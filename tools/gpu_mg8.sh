#!/bin/bash
# N GPUs (default 8): NCCL / push-exchange parity tests, the bench at N (strong scaling of the 8M workload), and at N = 8 the 32M pacing workload
N=${1:-8}
mkdir -p gpurun_out
if [ "$N" != "8" ]; then
timeout 900 python -m pytest tests/test_gpu_nccl.py -m gpu -q --timeout 600 > gpurun_out/pytest_nccl_$N.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_nccl_$N.log
grep -E "passed|failed|^FAILED|^E  |skipped" gpurun_out/pytest_nccl_$N.log | head -20
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_mg_$N.log 2>gpurun_out/bench_mg_$N.err; echo "bench $N exit $?"
grep "^{" gpurun_out/bench_mg_$N.log | python -c "
import sys,json
for ln in sys.stdin:
    d=json.loads(ln); print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e']['h2d_bytes_per_step'], d['e2e']['d2h_bytes_per_step'], d.get('mg_parity'), d['config'].get('exchange1','')[:12], d['config'].get('moment_allreduce','')[:12], d['roofline']['kernel_group_ms'])
"
if [ "$N" == "8" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus 8 --steps 100 --warmup 5 --workload 32m --no-cpu-baseline > gpurun_out/bench_mg_8_32m.log 2>&1; echo "bench 32m exit $?"
grep "^{" gpurun_out/bench_mg_8_32m.log | python -c "
import sys,json
for ln in sys.stdin:
    d=json.loads(ln); print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, 'e2e', d['e2e']['value'], d['roofline']['whole_step'])
"
fi
