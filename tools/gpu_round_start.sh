#!/bin/bash
# First GPU call of a round, in one gpurun invocation (saves box acquisitions):
#   gpurun --timeout 1500 -- 'bash tools/gpu_round_start.sh r02a'
# 1. the whole GPU suite (the tests/test_zz*_gpu_*.py files were written on the CPU only in round 1: their first device run),
# 2. both bench arms at N = 1, 3. the launch list of the bench command, 4. ncu --set full of the contraction kernel and of the
# tree-algebra kernels. Everything lands under gpurun_out/<tag>_*; copy what is to be judged into profiles/.
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
python -m pytest tests -m gpu -q -x --durations=15 > $out/${tag}_tests.txt 2>&1
# second pass without -x over the files that were new in round 1, so that one failure does not hide the others
python -m pytest tests/test_zz1_gpu_reference.py tests/test_zz3_gpu_cpp_mirror.py tests/test_zz2_gpu_tree_algebra.py -m gpu -q > $out/${tag}_tests_zz.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.txt 2>&1
python bench.py --impl reference > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/${tag}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_under_ncu.json 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:pipe_contract -s 20 -c 3 -o $out/${tag}_contract \
    python tools/scale_probe.py 1000 > /dev/null 2>&1
python tools/prof_algebra.py 100 7 1e-6 > $out/${tag}_algebra.txt 2>&1
ncu --set full --clock-control none --import-source on -k "regex:axpy_nodes|product_values|transform8|norms_kernel" -c 24 -o $out/${tag}_algebra \
    python -m pytest tests/test_zz2_gpu_tree_algebra.py -m gpu -q -k "multiply and 7" > /dev/null 2>&1
tail -3 $out/${tag}_tests.txt $out/${tag}_tests_zz.txt $out/${tag}_smoke.txt
cat $out/${tag}_bench_n1.json
