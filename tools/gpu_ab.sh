#!/bin/bash
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --e2e-steps 2 --cpu-sample 0 2>&1 | tail -1 | cut -c1-700
