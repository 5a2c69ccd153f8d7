#!/bin/bash
timeout 300 python tools/dbg_epi.py 2>&1 | grep "bad count"
timeout 300 python tools/dbg_epi.py 2>&1 | grep "bad count"
bash tools/gpu_run21.sh
