#!/bin/bash
# one GPU: the whole GPU test suite (virtual-rank slab tests included)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest.log | head -40
