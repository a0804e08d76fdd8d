# one gpurun call: the GPU test tier, smoke, the default bench line and the reference arm, the
# large-table timings
T=${1:-final4}
set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_gputest_$T.log 2>&1; tail -3 gpurun_out/r02_gputest_$T.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_$T.log 2>&1; tail -2 gpurun_out/r02_smoke_$T.log
python bench.py > gpurun_out/r02_bench_$T.json 2> gpurun_out/r02_bench_$T.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_${T}_reference_arm.json 2> gpurun_out/r02_bench_${T}_reference_arm.err; echo "reference arm rc=$?"
python tools/k1b_time.py > gpurun_out/r02_k1_large_tables_$T.log 2>&1; cat gpurun_out/r02_k1_large_tables_$T.log
