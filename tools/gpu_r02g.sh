#!/bin/bash
# round 2, GPU call g (1 GPU): headline bench with the new transforms rows; the same workload at 2000 / 3000 centres (N = 1 values of
# the scaling experiment); memory footprint
out=gpurun_out; tag=r02g; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python bench.py --centers 2000 --no-cpu-baseline > $out/${tag}_bench_n1_c2000.json 2> $out/${tag}_bench_n1_c2000.err
nvidia-smi --query-gpu=memory.used --format=csv >> $out/${tag}_host.txt
python bench.py --centers 3000 --no-cpu-baseline > $out/${tag}_bench_n1_c3000.json 2> $out/${tag}_bench_n1_c3000.err
free -g >> $out/${tag}_host.txt
python -c "
import json
for f in ('bench_n1','bench_n1_c2000','bench_n1_c3000'):
    try:
        d=json.load(open('$out/${tag}_'+f+'.json')); print(f, d['value'], d['ms_per_step'], d['e2e'], d['detail']['ms_not_in_kernels_per_step'], d['detail']['ms_contract_per_step'], d['detail']['output_nodes_per_step'], d['detail']['input_tree_nodes'], d['roofline']['frac'])
        if 'transforms' in d: print(json.dumps(d['transforms'])[:1500])
    except Exception as e: print(f, 'failed', e)
"
tail -3 $out/${tag}_bench_n1_c3000.err
