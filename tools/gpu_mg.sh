#!/bin/bash
# multi-GPU checks (run with gpurun --gpus N): NCCL slab parity vs single GPU, then the bench at N ranks
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/mg_parity.py --steps 30 --canonical > gpurun_out/mg_parity_$N.log 2>&1; echo "exit $?" >> gpurun_out/mg_parity_$N.log
tail -4 gpurun_out/mg_parity_$N.log
timeout 300 $TR tools/mg_parity.py --steps 30 --quadratic --dims 128x20x20 > gpurun_out/mg_parity_q_$N.log 2>&1; echo "exit $?" >> gpurun_out/mg_parity_q_$N.log
tail -3 gpurun_out/mg_parity_q_$N.log
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_mg_$N.log 2>&1; echo "exit $?" >> gpurun_out/bench_mg_$N.log
tail -3 gpurun_out/bench_mg_$N.log | cut -c1-1500
