#!/bin/bash
# round 2, GPU call o (1 GPU): persistent transform8 kernel (k = 7) + closing TopDown(+=) folded into the loop: whole GPU suite,
# C2 sweep with and without, headline bench with and without the fold
out=gpurun_out; tag=r02o; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
python -m pytest tests -m gpu -q -x > $out/${tag}_tests.txt 2>&1
python tools/prof_transform.py 1000 20 7 > $out/${tag}_transforms_pipe.txt 2>&1
MRX_NO_TPIPE=1 python tools/prof_transform.py 1000 20 7 > $out/${tag}_transforms_nopipe.txt 2>&1
python bench.py --no-cpu-baseline > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
MRX_NO_TDFOLD=1 python bench.py --no-cpu-baseline > $out/${tag}_bench_n1_nofold.json 2> $out/${tag}_bench_n1_nofold.err
tail -4 $out/${tag}_tests.txt; cat $out/${tag}_transforms_pipe.txt; echo ---; cat $out/${tag}_transforms_nopipe.txt
python -c "
import json
for f in ('bench_n1','bench_n1_nofold'):
    d=json.load(open('$out/${tag}_'+f+'.json')); print(f, d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['detail']['ms_post_per_step'], d['detail']['ms_not_in_kernels_per_step'])
d=json.load(open('$out/${tag}_bench_n1.json')); print(json.dumps(d['transforms'])[:1500])
"
