#!/bin/bash
mkdir -p gpurun_out/c18
O=gpurun_out/c18
timeout 900 python -m pytest tests/test_train_gpu.py -x -q -m gpu -s > $O/tests.log 2>&1; echo "exit $?" >> $O/tests.log
timeout 600 python -m pytest tests/test_parity_at_size_gpu.py -x -q -m gpu -k cfg4 -s 2>&1 | grep -E "cfg4|passed|failed" > $O/atsize.log
timeout 300 python bench.py --no-degrade --no-tiled --no-cpu --steps 10 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['value'], 'train', d['training'].get('ms_per_step'), d['training'].get('value'), d['training'].get('error'))
" > $O/train.log 2>&1
python tools/trace_train.py > $O/trace.txt 2>&1
grep -n "passed\|failed\|exit\|cosine" $O/tests.log | tail -n 12; cat $O/atsize.log $O/train.log; tail -n 22 $O/trace.txt
