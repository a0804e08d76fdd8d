# multi-GPU round, short form: weak config 2 and the sharded config 5 (gpurun --gpus N)
N=${1:-8}
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n${N}_config2_weak.json 2> gpurun_out/r02_bench_n${N}_config2_weak.err; echo "weak rc=$?"
$TR bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --scaling strong --workload config5 --strings $((N*1000000)) > gpurun_out/r02_bench_n${N}_config5_strong.json 2> gpurun_out/r02_bench_n${N}_config5_strong.err; echo "strong5 rc=$?"
