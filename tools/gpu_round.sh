set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_gputest6.log 2>&1; tail -3 gpurun_out/r02_gputest6.log
(python tools/mfa_time.py config3 k4 && python tools/mfa_time.py config5 k4) > gpurun_out/r02_mfa_time8.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r02_launches_c.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra-workloads > gpurun_out/r02_launches_c.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k4_mfa -c 1 -s 2 -o gpurun_out/r02_k4_v8_config3 python tools/mfa_time.py config3 k4 1000000 2 > gpurun_out/r02_k4_v8_ncu.log 2>&1; echo "ncu full rc=$?"
