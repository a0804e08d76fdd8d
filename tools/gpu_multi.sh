# multi-GPU round: N = number of GPUs of the box (gpurun --gpus N)
N=${1:-2}
set -x
python -m pytest tests/test_sharding_nccl_gpu.py -m gpu -q > gpurun_out/r02_nccl_test_n$N.log 2>&1; tail -2 gpurun_out/r02_nccl_test_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n${N}_config2_weak.json 2> gpurun_out/r02_bench_n${N}_config2_weak.err; echo "weak rc=$?"
$TR bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --scaling strong --workload config5 --strings $((N*1000000)) > gpurun_out/r02_bench_n${N}_config5_strong.json 2> gpurun_out/r02_bench_n${N}_config5_strong.err; echo "strong5 rc=$?"
$TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --scaling strong --workload config2 --strings 1000000 > gpurun_out/r02_bench_n${N}_config2_strong.json 2> gpurun_out/r02_bench_n${N}_config2_strong.err; echo "strong2 rc=$?"
