#!/bin/bash
# round 2, first GPU call: ncu evidence for the transform kernels and the k=9/11 contraction, per-iteration host profile
out=gpurun_out; tag=r02a; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
python tools/prof_transform.py 100 20 5 7 9 11 > $out/${tag}_transforms.txt 2>&1
ncu --set full --clock-control none --import-source on -k "regex:transform8_kernel|transformK_kernel" -c 60 -o $out/${tag}_transforms \
    python tools/prof_transform.py 100 1 7 9 > /dev/null 2>&1
MRX_PROBE_K=9 ncu --set full --clock-control none --import-source on -k regex:pipe_contract_coop -s 12 -c 2 -o $out/${tag}_coop10 \
    python tools/scale_probe.py 100 > /dev/null 2>&1
MRX_PROBE_K=11 MRX_PROBE_PREC=1e-9 ncu --set full --clock-control none --import-source on -k regex:pipe_contract_coop -s 12 -c 2 -o $out/${tag}_coop12 \
    python tools/scale_probe.py 30 > /dev/null 2>&1
MRX_PROFILE=1 python tools/scale_probe.py 1000 > $out/${tag}_profile1000.txt 2>&1
MRX_PROBE_K=9 python tools/scale_probe.py 100 > $out/${tag}_k9.txt 2>&1
MRX_PROBE_K=11 MRX_PROBE_PREC=1e-9 python tools/scale_probe.py 100 > $out/${tag}_k11.txt 2>&1
cat $out/${tag}_transforms.txt; tail -5 $out/${tag}_k9.txt $out/${tag}_k11.txt; tail -40 $out/${tag}_profile1000.txt
