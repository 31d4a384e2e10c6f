set -x
cd $GRAFT_REPO_ROOT
AGB_DEBUG_PLAN=1 timeout 300 python tools/one_scan.py 64 "because each" k=2 list=1 reps=4 > gpurun_out/r2i_one64.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2i_launches_64GiB.csv python tools/one_scan.py 64 "because each" k=2 list=1 reps=2 > gpurun_out/r2i_ncu_launches.log 2>&1
cat gpurun_out/r2i_one64.log
grep -v "^==" gpurun_out/r2i_launches_64GiB.csv | tail -10 | awk -F'","' '{print substr($5,1,50), $NF}'
