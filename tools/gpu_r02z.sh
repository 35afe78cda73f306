#!/bin/bash
# round 2, GPU call z (1 GPU): waiting contraction kernel with weak block loads (call y: __ldca had compiled to LDG.STRONG.SM);
# clean A/B of the end-to-end arm with and without the gather beside the contraction
out=gpurun_out; tag=r02z; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mirror or host_memory or end_to_end or full_size" > $out/${tag}_tests.txt 2>&1
tail -3 $out/${tag}_tests.txt
MRX_E2E_MIRROR_ONLY=1 MRX_E2E_KEEP=1 timeout 100 python tools/prof_e2e.py 1000 > $out/${tag}_e2e_default.txt 2>&1
echo "default: $(grep 'mirror True\|partially' $out/${tag}_e2e_default.txt | tail -4 | tr '\n' ' ')"
MRX_NO_FETCH_OVERLAP=1 MRX_E2E_MIRROR_ONLY=1 MRX_E2E_KEEP=1 timeout 100 python tools/prof_e2e.py 1000 > $out/${tag}_e2e_nooverlap.txt 2>&1
echo "nooverlap: $(grep 'mirror True\|partially' $out/${tag}_e2e_nooverlap.txt | tail -4 | tr '\n' ' ')"
timeout 300 python bench.py --no-cpu-baseline > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
MRX_NO_FETCH_OVERLAP=1 timeout 300 python bench.py --no-cpu-baseline > $out/${tag}_bench_n1_nooverlap.json 2> $out/${tag}_bench_n1_nooverlap.err
MRX_NO_FETCH_OVERLAP=1 MRX_SUB_RANGES=1 timeout 300 python bench.py --no-cpu-baseline > $out/${tag}_bench_n1_neither.json 2> $out/${tag}_bench_n1_neither.err
python -c "
import json
for f in ('','_nooverlap','_neither'):
    d=json.load(open('$out/${tag}_bench_n1'+f+'.json')); print(f, d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'])
"
