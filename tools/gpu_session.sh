set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_shard_nccl.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_nccl_tests.log
cat gpurun_out/r2_nccl_tests.log
for N in 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2950$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
cat gpurun_out/r2_bench_n$N.json; tail -3 gpurun_out/r2_bench_n$N.err
done
