#!/bin/bash
# round 2, GPU call ae (1 GPU), last of the round: whole GPU suite + smoke + bench (headline, c5, c4) on the code as committed, and
# the ncu launch list of the bench command (shares of the step)
out=gpurun_out; tag=r02ae; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
timeout 220 python -m pytest tests -m gpu -q -x > $out/${tag}_tests.txt 2>&1
tail -2 $out/${tag}_tests.txt
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/${tag}_smoke.txt 2>&1
tail -1 $out/${tag}_smoke.txt
timeout 150 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
timeout 60 python bench.py --config c5 --steps 3 --no-cpu-baseline > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err
timeout 60 python bench.py --config c4 --steps 2 --no-cpu-baseline > $out/${tag}_bench_c4.json 2> $out/${tag}_bench_c4.err
python -c "
import json
for f in ('bench_n1','bench_c5','bench_c4'):
    try:
        d=json.load(open('$out/${tag}_'+f+'.json')); print(f, d['value'], d['ms_per_step'], d.get('e2e'), (d.get('roofline') or {}).get('frac'), d.get('breakdown_ms'))
    except Exception as e: print(f, 'failed', e)
"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $out/${tag}_ncu.log 2>&1
python tools/launch_shares.py $out/${tag}_launches.csv 6 > $out/${tag}_launch_shares.csv 2>&1
rm -f $out/${tag}_launches.csv
head -12 $out/${tag}_launch_shares.csv
