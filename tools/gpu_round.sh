# one gpurun call: the GPU test tier, the default bench line, its launch list, ncu captures of the two headline kernels
T=${1:-final}
set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_gputest_$T.log 2>&1; tail -3 gpurun_out/r02_gputest_$T.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_$T.log 2>&1; tail -2 gpurun_out/r02_smoke_$T.log
python bench.py > gpurun_out/r02_bench_$T.json 2> gpurun_out/r02_bench_$T.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_${T}_reference_arm.json 2> gpurun_out/r02_bench_${T}_reference_arm.err; echo "reference arm rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r02_launches_$T.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra-workloads > gpurun_out/r02_launches_$T.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k4_mfa -c 1 -s 2 -o gpurun_out/r02_k4_${T}_config3 python tools/mfa_time.py config3 k4 1000000 2 > gpurun_out/r02_k4_${T}_ncu.log 2>&1; echo "ncu k4 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k1_rows|k1_tilesort' -c 2 -s 6 -o gpurun_out/r02_k1_$T python tools/k1_time.py 1000000 3 > gpurun_out/r02_k1_${T}_ncu.log 2>&1; echo "ncu k1 rc=$?"
(python tools/mfa_time.py config3 k4 && python tools/mfa_time.py config5 k4 && python tools/mfa_time.py config4 k4) > gpurun_out/r02_mfa_time_$T.log 2>&1
