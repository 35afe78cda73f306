#!/bin/bash
# round 2, GPU call j (1 GPU): host mirror with contiguous per-chunk copies: parity + e2e bench
out=gpurun_out; tag=r02j; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "host_mirror or host_memory" > $out/${tag}_tests_new.txt 2>&1
python bench.py --no-cpu-baseline > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
tail -3 $out/${tag}_tests_new.txt
python -c "
import json
for f in ('bench_n1',):
    d=json.load(open('$out/${tag}_'+f+'.json')); print(f, d['value'], d['ms_per_step'], d['e2e'], d['same_workload_as_reference_arm']['e2e_ms_per_step'], d['same_workload_as_reference_arm']['ms_per_step'])
"
