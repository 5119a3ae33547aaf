set -x
mkdir -p gpurun_out/r01f
python bench.py > gpurun_out/r01f/bench_n1.json 2> gpurun_out/r01f/bench_n1.err; echo rc=$?
python bench.py --impl reference > gpurun_out/r01f/bench_reference_arm.json 2> gpurun_out/r01f/ref.err; echo rc=$?
python bench.py --steps 3 --warmup 3 --docs 50000 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r01f/small.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01f/launches_bench_docs50000.csv \
  python bench.py --steps 3 --warmup 3 --docs 50000 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r01f/ncu_list.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:lda_md_int_sweep_half16 -s 3 -c 1 -o gpurun_out/r01f/prof_h16 -f \
  python bench.py --steps 3 --warmup 3 --docs 50000 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r01f/ncu_full.log 2>&1; echo rc=$?
ls -la gpurun_out/r01f
