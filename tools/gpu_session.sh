set -x
cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests/test_gpu_dropin.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2z_tests.log
grep -E "^E|passed|failed" gpurun_out/r2z_tests.log | head
