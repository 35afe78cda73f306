#!/bin/bash
# round 2, GPU call q (1 GPU): closing mirror push limited to branch nodes; latency profile of a small workload (125 centres at one GPU
# has about the per-rank kernel time of 1000 centres at 8) with its launch list
out=gpurun_out; tag=r02q; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mirror or host_resident or lazy" > $out/${tag}_tests.txt 2>&1
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
MRX_PROFILE=1 python tools/scale_probe.py 125 > $out/${tag}_profile_125.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches_125.csv python tools/scale_probe.py 125 > $out/${tag}_ncu_125.log 2>&1
python tools/launch_shares.py $out/${tag}_launches_125.csv 6 > $out/${tag}_launch_shares_125.csv 2>&1
rm -f $out/${tag}_launches_125.csv
tail -3 $out/${tag}_tests.txt
grep "rank 0\|device_apply ms\|host phases\|run_apply_pipe" $out/${tag}_profile_125.txt | tail -8
head -30 $out/${tag}_launch_shares_125.csv
python -c "
import json
d=json.load(open('$out/${tag}_bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['breakdown_ms'])
"
