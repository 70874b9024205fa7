#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 200 python tools/perf_probe.py 5 70 20 4096 200 fp16 1 0 0 2>&1 | tail -2 | cut -c1-330
