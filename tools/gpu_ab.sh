#!/bin/bash
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --gpus 1 --steps 2 --warmup 3 --e2e-steps 3 --cpu-sample 0 2>&1 | tail -1 | cut -c1-1200
