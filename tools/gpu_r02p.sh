#!/bin/bash
# round 2, GPU call p (2 GPUs): shared host arena (every rank downloads its share of the result): sharded tests incl. the shared mirror,
# periodic C++ program on the device, bench N = 2 with and without the shared arena
out=gpurun_out; tag=r02p; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
df -h /dev/shm > $out/${tag}_shm.txt 2>&1
MRX_EXPECT_ARENA=1 python -m pytest tests/test_gpu_sharded.py tests/test_zz3_gpu_cpp_mirror.py -m gpu -q -x > $out/${tag}_tests.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29531 bench.py --gpus 2 > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.err
MRX_BENCH_NO_SHARED_MIRROR=1 $TR --master-port 29532 bench.py --gpus 2 > $out/${tag}_bench_n2_rank0.json 2> $out/${tag}_bench_n2_rank0.err
tail -15 $out/${tag}_tests.txt
cat $out/${tag}_shm.txt
python -c "
import json
for f in ('bench_n2','bench_n2_rank0'):
    try:
        d=json.load(open('$out/${tag}_'+f+'.json')); print(f, d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['breakdown_ms'])
    except Exception as e: print(f, 'failed', e)
"
tail -5 $out/${tag}_bench_n2.err
