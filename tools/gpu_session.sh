set -x
cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2final2_tests.log
grep -E "^E|passed|failed" gpurun_out/r2final2_tests.log | head
timeout 300 python tools/one_scan.py 64 "because each" k=2 list=1 linenum=1 ordinals=1 reps=3 2>&1 | tail -2
