#!/bin/bash
# round 2, GPU call ak (1 GPU), the last seconds of the budget: the bench line with the per-call conversion path (from_caller_arrays)
out=gpurun_out; tag=r02ak; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
timeout 50 python bench.py --steps 2 --no-cpu-baseline > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python -c "
import json
d=json.load(open('$out/${tag}_bench_n1.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['same_workload_as_reference_arm'])
"
tail -2 $out/${tag}_bench_n1.err
