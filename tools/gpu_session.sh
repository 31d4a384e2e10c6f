set -x
cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2m_tests.log
tail -8 gpurun_out/r2m_tests.log
timeout 300 python tools/one_scan.py 64 "the" reps=3 > gpurun_out/r2m_the64.log 2>&1
cat gpurun_out/r2m_the64.log
