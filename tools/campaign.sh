#!/bin/bash
# Measurement campaign on one GPU (run under gpurun): bench line, reference arm, ncu launch list of the bench
# command at 50,000 documents, one `ncu --set full` capture of the headline kernel.  Outputs in gpurun_out/<tag>/.
tag=${1:-campaign}
out=gpurun_out/$tag
mkdir -p $out
python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; echo bench rc=$?
python bench.py --impl reference > $out/bench_reference_arm.json 2> $out/ref.err; echo reference rc=$?
small="python bench.py --steps 3 --warmup 3 --docs 50000 --no-cpu-baseline --e2e-steps 1"
$small > $out/small.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_bench_docs50000.csv $small > $out/ncu_list.log 2>&1; echo launch-list rc=$?
ncu --set full --clock-control none --import-source on -k regex:lda_md_int_sweep_half16 -s 3 -c 1 -o $out/prof_h16 -f $small > $out/ncu_full.log 2>&1; echo full rc=$?
