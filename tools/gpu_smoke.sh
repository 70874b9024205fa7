#!/bin/bash
# smallest end-to-end check on a fresh box: smoke() (which calls build()) and the reference arm of the bench
mkdir -p gpurun_out
ls -la nested_diffusion_b200/lib/
echo "== smoke"; ( time timeout 300 python __graft_entry__.py smoke ) 2>&1 | tail -8
echo "== bench --impl reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-700
