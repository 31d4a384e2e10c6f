set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2v_tests.log
grep -E "^E|passed|failed" gpurun_out/r2v_tests.log | head
timeout 600 python tools/path_bench.py > gpurun_out/r2v_path_bench.log 2>&1
tail -17 gpurun_out/r2v_path_bench.log
