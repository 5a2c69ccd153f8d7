#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py -m gpu -x -q -k epilogues > gpurun_out/conv.log 2>&1; grep -E "^E  |assert" gpurun_out/conv.log | head -20
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; tail -c 3000 gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2
