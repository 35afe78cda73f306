#!/bin/bash
# round 2, GPU call r (2 GPUs): block-cyclic dealing of the work vector (MRX_SHARD_BLOCK = 8, 64 against plain cyclic): bit-identity,
# resident and end-to-end step, bytes gathered per rank
out=gpurun_out; tag=r02r; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
MRX_SHARD_BLOCK=8 MRX_EXPECT_ARENA=1 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x > $out/${tag}_tests_b8.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for B in 8 64 1; do
MRX_SHARD_BLOCK=$B $TR --master-port 2953$((B % 10)) bench.py --gpus 2 > $out/${tag}_bench_n2_b$B.json 2> $out/${tag}_bench_n2_b$B.err
done
tail -3 $out/${tag}_tests_b8.txt
python -c "
import json
for f in ('b8','b64','b1'):
    try:
        d=json.load(open('$out/${tag}_bench_n2_'+f+'.json')); print(f, d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['breakdown_ms'])
    except Exception as e: print(f, 'failed', e)
"
tail -5 $out/${tag}_bench_n2_b8.err
