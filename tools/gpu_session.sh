set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_shard.py tests/test_gpu_shard_nccl.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2g_tests.log
tail -25 gpurun_out/r2g_tests.log
AGB_BENCH_SECONDARY=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err
tail -5 gpurun_out/r2g_bench_n2.err
cut -c1-900 gpurun_out/r2g_bench_n2.json
