set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2e_tests.log
timeout 300 python tools/one_scan.py 64 "because each" k=2 list=1 reps=4 > gpurun_out/r2e_one64.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2e_launches_64GiB.csv python tools/one_scan.py 64 "because each" k=2 list=1 reps=2 > gpurun_out/r2e_ncu_launches.log 2>&1
tail -3 gpurun_out/r2e_tests.log
cat gpurun_out/r2e_one64.log
grep -v "^==" gpurun_out/r2e_launches_64GiB.csv | tail -9 | awk -F'","' '{print substr($5,1,50), $NF}'
