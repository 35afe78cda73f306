#!/bin/bash
# round 2, GPU call i (1 GPU): host mirror (streamed result download) + block-granular input gather: parity tests, e2e bench, binding
out=gpurun_out; tag=r02i; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
python -m pytest tests/test_gpu_parity.py tests/test_ref_binding.py -m gpu -q -x -k "host_mirror or host_memory or c_abi_on_the_device" -rP > $out/${tag}_tests_new.txt 2>&1
python bench.py --no-cpu-baseline > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
MRX_NO_MIRROR_STREAM=1 python bench.py --no-cpu-baseline > $out/${tag}_bench_n1_nomirror.json 2> $out/${tag}_bench_n1_nomirror.err
python -m pytest tests -m gpu -q -x > $out/${tag}_tests.txt 2>&1
grep -v "^\[parity\]" $out/${tag}_tests_new.txt | tail -15; tail -4 $out/${tag}_tests.txt
python -c "
import json
for f in ('bench_n1','bench_n1_nomirror'):
    d=json.load(open('$out/${tag}_'+f+'.json')); print(f, d['value'], d['ms_per_step'], d['e2e'], d['same_workload_as_reference_arm']['e2e_ms_per_step'], d['same_workload_as_reference_arm']['ms_per_step'])
"
