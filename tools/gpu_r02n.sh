#!/bin/bash
# round 2, GPU call n (2 GPUs): sharded apply after the loop changes (mapped flags arena, mirror on rank 0, block gather): bit-identity
# tests + bench N = 2 (headline and c5) + N = 1 on the same box
out=gpurun_out; tag=r02n; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
python -m pytest tests/test_gpu_sharded.py -m gpu -q -x > $out/${tag}_tests_sharded.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29531 bench.py --gpus 2 > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.err
$TR --master-port 29532 bench.py --gpus 2 --config c5 --steps 3 > $out/${tag}_bench_n2_c5.json 2> $out/${tag}_bench_n2_c5.err
$TR --master-port 29533 bench.py --gpus 2 --config c4 --steps 2 > $out/${tag}_bench_n2_c4.json 2> $out/${tag}_bench_n2_c4.err
tail -3 $out/${tag}_tests_sharded.txt
python -c "
import json
for f in ('bench_n2','bench_n2_c5','bench_n2_c4'):
    try:
        d=json.load(open('$out/${tag}_'+f+'.json')); print(f, d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'])
    except Exception as e: print(f, 'failed', e)
"
tail -3 $out/${tag}_bench_n2_c4.err
