set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_dropin.py -x -q -m gpu -k three_gib 2>&1 | tail -60 > gpurun_out/r2u_tests.log
grep -E "^E|assert|passed|failed" gpurun_out/r2u_tests.log | head -30
timeout 600 python tools/one_scan.py 64 "because each" k=2 list=1 linenum=1 ordinals=1 reps=3 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_shard.py -x -q -m gpu 2>&1 | tail -3
