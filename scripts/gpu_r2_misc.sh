#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_reference_suite.py -m gpu -x -q -s 2>&1 | tail -60
bash scripts/gpu_sanitize_r2.sh
