#!/bin/bash
# round 2, GPU call b: full GPU suite (new parity cases at BASELINE's stated configs), bench lines (headline, reference arm, c4, c5),
# ncu --set full of the transform kernels and of the k=9 / k=11 contraction, exported to CSV on the box (reports stay small)
out=gpurun_out; tag=r02b; mkdir -p $out
free -g > $out/${tag}_host.txt; nproc >> $out/${tag}_host.txt; nvidia-smi -L >> $out/${tag}_host.txt
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
python -m pytest tests -m gpu -q -x --durations=20 -rP -k "baseline_configs or zz1" > $out/${tag}_tests_new.txt 2>&1
python -m pytest tests -m gpu -q --durations=10 > $out/${tag}_tests.txt 2>&1
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python bench.py --impl reference > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
python bench.py --config c5 --steps 3 > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err
python bench.py --config c4 --steps 3 > $out/${tag}_bench_c4.json 2> $out/${tag}_bench_c4.err
python tools/prof_transform.py 1000 20 5 7 9 > $out/${tag}_transforms.txt 2>&1
cap() { # name, kernel regex, skip, count, command...
  name=$1; rx=$2; skip=$3; cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $cnt -f -o $out/${tag}_$name "$@" > /dev/null 2>&1
  ncu -i $out/${tag}_$name.ncu-rep --page raw --csv > $out/${tag}_${name}_raw.csv 2>/dev/null
  ncu -i $out/${tag}_$name.ncu-rep --page details --csv > $out/${tag}_${name}_details.csv 2>/dev/null
  sz=$(stat -c %s $out/${tag}_$name.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 12000000 ]; then rm -f $out/${tag}_$name.ncu-rep; fi
}
cap tf7 "transform8_kernel" 0 40 python tools/prof_transform.py 1000 1 7
cap tf9 "transformK_kernel" 0 40 python tools/prof_transform.py 400 1 9
cap tf5 "transformK_kernel" 0 40 python tools/prof_transform.py 1000 1 5
MRX_PROBE_K=9 cap coop10 "pipe_contract_coop" 12 2 python tools/scale_probe.py 100
MRX_PROBE_K=11 MRX_PROBE_PREC=1e-9 cap coop12 "pipe_contract_coop" 12 2 python tools/scale_probe.py 30
du -sh $out; ls -la $out
tail -5 $out/${tag}_tests_new.txt; tail -5 $out/${tag}_tests.txt; cat $out/${tag}_transforms.txt
