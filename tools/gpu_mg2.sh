#!/bin/bash
# N GPUs: NCCL parity tests, then the bench at N (and the host-side profile of the slab step)
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nccl.py -m gpu -q --timeout 600 > gpurun_out/pytest_nccl.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_nccl.log
grep -E "passed|failed|^FAILED|^E  |skipped" gpurun_out/pytest_nccl.log | head -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_mg_$N.log 2>gpurun_out/bench_mg_$N.err; echo "bench exit $?"
grep "^{" gpurun_out/bench_mg_$N.log | python -c "
import sys,json
for ln in sys.stdin:
    d=json.loads(ln); print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, d['e2e']['value'], d['e2e']['h2d_bytes_per_step'], d.get('mg_parity'), d['roofline']['kernel_group_ms'])
"
tail -3 gpurun_out/bench_mg_$N.err
# the sized ncclSend / ncclRecv messages (opt-in) on a jittered lattice
SPHSM_P2P=0 SPHSM_X1_DYNAMIC=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 tools/mg_parity.py --steps 30 --quadratic --dims 128x20x20 2>/dev/null | grep "^{"
SPHSM_HOST_PROF=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 128 --warmup 5 --no-cpu-baseline > gpurun_out/hostprof_$N.log 2>&1; grep "sphsm" gpurun_out/hostprof_$N.log | tail -4
