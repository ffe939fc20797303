#!/bin/bash
mkdir -p gpurun_out
python tests/golden/gen_engine_draw_goldens.py --dump 2>&1 | tail -2
timeout 500 python -m pytest tests -m gpu -x -q -k "not identical_draws" 2>&1 | tail -8
