#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mgpu.py -x -q -m gpu > gpurun_out/mgpu_pytest.log 2>&1
grep -n "rank[0-9]\]:\|Assertion\|assert \|case [0-9]" gpurun_out/mgpu_pytest.log | cut -c1-400 | head -40
tail -3 gpurun_out/mgpu_pytest.log
