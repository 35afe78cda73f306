#!/bin/bash
# round 2, GPU call s (1 GPU): whole GPU suite + smoke + bench (all configs) on the code as committed; staged uploads of norms/topology
out=gpurun_out; tag=r02s; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
python -m pytest tests -m gpu -q -x > $out/${tag}_tests.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/${tag}_smoke.txt 2>&1
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python bench.py --impl reference > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
python bench.py --config c5 --steps 3 --no-cpu-baseline > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err
python bench.py --config c4 --steps 2 --no-cpu-baseline > $out/${tag}_bench_c4.json 2> $out/${tag}_bench_c4.err
python bench.py --config c1 --no-cpu-baseline > $out/${tag}_bench_c1.json 2> $out/${tag}_bench_c1.err
MRX_PROFILE=1 python tools/prof_e2e.py 1000 > $out/${tag}_e2e_phases.txt 2>&1
tail -4 $out/${tag}_tests.txt; tail -2 $out/${tag}_smoke.txt
python -c "
import json
for f in ('bench_n1','bench_c5','bench_c4','bench_c1','bench_ref'):
    try:
        d=json.load(open('$out/${tag}_'+f+'.json')); print(f, d['value'], d['ms_per_step'], d.get('e2e'), (d.get('roofline') or {}).get('frac'), d.get('breakdown_ms'))
    except Exception as e: print(f, 'failed', e)
"
grep "mirror True rep 2\|mirror False rep 2" $out/${tag}_e2e_phases.txt | head -3
grep "device_apply ms\|run_apply_pipe ms\|push of the" $out/${tag}_e2e_phases.txt | sed -n 10,16p
