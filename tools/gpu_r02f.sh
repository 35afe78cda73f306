#!/bin/bash
# round 2, GPU call f: periodic apply / near / far field on the device; K = 10 stack variants; regression of the whole suite
out=gpurun_out; tag=r02f; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
python -m pytest tests/test_gpu_periodic.py -m gpu -q --durations=5 > $out/${tag}_tests_periodic.txt 2>&1
for v in 0 1 2 3; do
  MRX_STACK_VARIANT=$v python bench.py --config c4 --steps 2 --no-cpu-baseline > $out/${tag}_bench_c4_v$v.json 2> $out/${tag}_bench_c4_v$v.err
done
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_periodic.py > $out/${tag}_tests.txt 2>&1
tail -25 $out/${tag}_tests_periodic.txt; tail -4 $out/${tag}_tests.txt
python -c "
import json
for v in range(4):
    d=json.load(open('$out/${tag}_bench_c4_v%d.json'%v)); print('variant',v, d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'])
"
