set -x
cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2n_tests.log
tail -8 gpurun_out/r2n_tests.log
timeout 300 python tools/one_scan.py 64 "because each" k=2 list=1 reps=4 > gpurun_out/r2n_one64.log 2>&1
cat gpurun_out/r2n_one64.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2n_launches_64GiB.csv python tools/one_scan.py 64 "because each" k=2 list=1 reps=2 > gpurun_out/r2n_ncu_launches.log 2>&1
grep -v "^==" gpurun_out/r2n_launches_64GiB.csv | tail -9 | awk -F'","' '{print substr($5,1,50), $NF}'
