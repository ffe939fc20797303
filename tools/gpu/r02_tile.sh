#!/bin/bash
for T in 1024 2048 4096; do echo "== T=$T"; CWQ_FRONT_TILE=$T timeout 200 python tools/profile_small.py --batch | cut -c1-420; done
for T in 1024 2048; do echo "== T=$T small"; CWQ_FRONT_TILE=$T timeout 200 python tools/profile_small.py | cut -c1-330; done
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
