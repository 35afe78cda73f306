#!/bin/bash
# round 2, GPU call ac (4 GPUs): sharded apply with the sub-range gather (host-resident input) and the shared host arena with more
# than two ranks: bit-identity tests, bench at 4 GPUs (resident and end to end)
out=gpurun_out; tag=r02ac; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
MRX_EXPECT_ARENA=1 timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x > $out/${tag}_tests.txt 2>&1
tail -3 $out/${tag}_tests.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29541 bench.py --gpus 4 --no-cpu-baseline > $out/${tag}_bench_n4.json 2> $out/${tag}_bench_n4.err
python -c "
import json
for f in ('bench_n4',):
    try:
        d=json.load(open('$out/${tag}_'+f+'.json')); print(f, d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['breakdown_ms'])
    except Exception as e: print(f, 'failed', e)
"
tail -3 $out/${tag}_bench_n4.err
