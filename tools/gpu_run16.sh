#!/bin/bash
for m in 30 14 22 6; do echo "mode $m"; RESR_WG_DBG=$m RESR_DEBUG_SYNC=1 timeout 60 python tools/wg_dbg.py 2>&1 | grep -E "resr\] wgrad:|rc" | tail -2; done
cuobjdump -sass real_esrgan-pytorch_b200/lib/libresr.so | grep -B3 -A3 "UTMALDG" | grep -A40 "wgrad" | head -5
