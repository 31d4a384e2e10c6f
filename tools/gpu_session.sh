set -x
cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2x_tests.log
grep -E "^E|passed|failed" gpurun_out/r2x_tests.log | head
timeout 600 python tools/path_bench.py > gpurun_out/r2x_path_bench.log 2>&1
tail -6 gpurun_out/r2x_path_bench.log
