set -x
python tools/k1_time.py 1000000 20 > gpurun_out/r02_k1_product_time.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r02_launches_b.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launches_b.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k1_rows|k1_tilesort' -c 2 -s 6 -o gpurun_out/r02_k1_rows_v2 python tools/k1_time.py 1000000 3 > gpurun_out/r02_k1_rows_v2_ncu.log 2>&1; echo "ncu full rc=$?"
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_sanitizer_memcheck.log python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "k1 or every_length or quad or golden" > gpurun_out/r02_sanitizer_memcheck_pytest.log 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/r02_sanitizer_memcheck_pytest.log; tail -5 gpurun_out/r02_sanitizer_memcheck.log
