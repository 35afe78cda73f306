#!/bin/bash
# round 2, GPU call m (1 GPU): split flags through the mapped arena (no copy-engine transfer in the loop): e2e with the mirror
out=gpurun_out; tag=r02m; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
python tools/prof_e2e.py 1000 > $out/${tag}_e2e.txt 2>&1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -q -x > $out/${tag}_tests.txt 2>&1
python bench.py --no-cpu-baseline > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
tail -3 $out/${tag}_tests.txt; grep "rep 2" $out/${tag}_e2e.txt
python -c "
import json
d=json.load(open('$out/${tag}_bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['same_workload_as_reference_arm']['e2e_ms_per_step'], d['same_workload_as_reference_arm']['ms_per_step'])
"
