set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "inverse_count or random_metachar or ragged" 2>&1 | tail -15 > gpurun_out/r2c_tests.log
grep -E "^E|passed|failed" gpurun_out/r2c_tests.log | head
timeout 200 python tools/path_bench.py 2>&1 | grep -E "inverse|pattern " 
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2final3_tests.log
grep -E "^E|passed|failed" gpurun_out/r2final3_tests.log | head
