#!/bin/bash
# round 2, GPU call ab (1 GPU): call aa had the node-range filter of the fill pass in the screen kernel (every candidate dropped);
# gather of sub-range s + 1 beside the contraction of sub-range s: tests, A/B of the end-to-end arm
out=gpurun_out; tag=r02ab; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mirror or host_memory or end_to_end or full_size" > $out/${tag}_tests.txt 2>&1
tail -3 $out/${tag}_tests.txt
timeout 150 python bench.py --no-cpu-baseline > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
MRX_NO_FETCH_OVERLAP=1 timeout 150 python bench.py --no-cpu-baseline > $out/${tag}_bench_n1_nooverlap.json 2> $out/${tag}_bench_n1_nooverlap.err
MRX_SUB_RANGES=7 MRX_SUB_MIN_TILES=2 timeout 150 python bench.py --no-cpu-baseline > $out/${tag}_bench_n1_sub7.json 2> $out/${tag}_bench_n1_sub7.err
python -c "
import json
for f in ('','_nooverlap','_sub7'):
    try:
        d=json.load(open('$out/${tag}_bench_n1'+f+'.json')); print(f, d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'])
    except Exception as e: print(f, 'failed', e)
"
