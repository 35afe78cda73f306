#!/bin/bash
# round 2, GPU call ad (8 GPUs): the bench line at 8 GPUs on the code as committed (resident and end to end: shared host arena with 8 ranks)
out=gpurun_out; tag=r02ad; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29551 bench.py --gpus 8 --no-cpu-baseline > $out/${tag}_bench_n8.json 2> $out/${tag}_bench_n8.err
python -c "
import json
for f in ('bench_n8',):
    try:
        d=json.load(open('$out/${tag}_'+f+'.json')); print(f, d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['breakdown_ms'])
    except Exception as e: print(f, 'failed', e)
"
tail -3 $out/${tag}_bench_n8.err
