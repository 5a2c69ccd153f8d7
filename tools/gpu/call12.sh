#!/bin/bash
mkdir -p gpurun_out/c12
O=gpurun_out/c12
timeout 300 python -m pytest tests/test_generator_gpu.py -q -m gpu -k "second_device" > $O/two_dev.log 2>&1; echo "exit $?" >> $O/two_dev.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "exit $?" >> $O/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/ref_n2.json 2> $O/ref_n2.err
python tools/time_degrade_ops.py > $O/ops.txt 2>&1
tail -n 3 $O/two_dev.log; tail -n 5 $O/bench_n2.err; head -c 1500 $O/bench_n2.json; echo; head -c 300 $O/ref_n2.json; echo; cat $O/ops.txt | tail -n 13
