#!/bin/bash
# one GPU: the whole GPU test suite, then the default bench line (what the driver runs), then the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest.log | head -20
cp gpurun_out/parity_report.json gpurun_out/parity_r02.json 2>/dev/null
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1.log 2>gpurun_out/bench_n1.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_n1.log") if l.startswith("{")][-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["roofline"]["kernel_group_ms"])
print("also", {k:(v["ms_per_step"], v["value"]) for k,v in d.get("also",{}).items()})
print("cpu", d.get("cpu_baseline"))
PY
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.log 2>&1; tail -c 600 gpurun_out/bench_ref.log
