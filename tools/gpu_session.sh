set -x
cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2s_tests.log
tail -12 gpurun_out/r2s_tests.log
