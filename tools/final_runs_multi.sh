# multi-GPU evidence of the final build: bash tools/final_runs_multi.sh N  -> gpurun_out/final/
N=$1; O=gpurun_out/final; mkdir -p $O
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; }
run --steps 30 --warmup 5 > $O/r02_bench_${N}gpu_weak.json 2> $O/weak$N.err
if [ $N -ge 4 ]; then run --steps 30 --warmup 5 --scaling strong --global-pairs 64 --no-e2e > $O/r02_bench_${N}gpu_strong.json 2> $O/strong$N.err; fi
if [ $N -eq 8 ]; then
  run --steps 20 --warmup 3 --workload cfg5 > $O/r02_bench_8gpu_cfg5.json 2> $O/cfg5.err
  timeout 600 python -m pytest tests/test_nccl_gpu.py -m gpu -q 2>&1 | tail -3 > $O/nccl_test.log
fi
tail -c 400 $O/r02_bench_${N}gpu_weak.json
