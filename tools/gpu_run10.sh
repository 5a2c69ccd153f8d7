#!/bin/bash
for n in 4 8 12 16 24 32 64; do timeout 300 python tools/time_generator.py $n 128 128 2>&1 | tail -1; done
