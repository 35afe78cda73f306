#!/bin/bash
# round 2, GPU call c: mailbox read-backs + pipelined split kernel; precTrees on the device; node transforms; the reference-side
# binding; baseline-config parity. Lean: new tests first (-x), then the full suite, bench headline, one profile run.
out=gpurun_out; tag=r02c; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
python -m pytest tests/test_gpu_prec_trees.py tests/test_gpu_node_transforms.py tests/test_ref_binding.py tests/test_gpu_baseline_configs.py -m gpu -q --durations=10 -rP > $out/${tag}_tests_new.txt 2>&1
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_baseline_configs.py --deselect tests/test_gpu_prec_trees.py --deselect tests/test_gpu_node_transforms.py --deselect tests/test_ref_binding.py > $out/${tag}_tests.txt 2>&1
python bench.py --no-cpu-baseline > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
MRX_NO_MAILBOX=1 python bench.py --no-cpu-baseline > $out/${tag}_bench_n1_nomailbox.json 2> $out/${tag}_bench_n1_nomailbox.err
MRX_PROFILE=1 python tools/scale_probe.py 1000 > $out/${tag}_profile1000.txt 2>&1
grep -v "^\[parity\]" $out/${tag}_tests_new.txt | tail -40; tail -5 $out/${tag}_tests.txt; tail -12 $out/${tag}_profile1000.txt
python -c "
import json
for f in ('${tag}_bench_n1.json','${tag}_bench_n1_nomailbox.json'):
    d=json.load(open('$out/'+f)); print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d['detail']['ms_not_in_kernels_per_step'], d['detail']['ms_kernel_per_step'])
"
