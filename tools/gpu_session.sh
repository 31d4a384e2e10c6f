set -x
cd $GRAFT_REPO_ROOT
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2_launches_bench_64GiB.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2_bench_under_ncu.log 2>&1
grep -v "^==" gpurun_out/r2_launches_bench_64GiB.csv | tail -12 | awk -F'","' '{print substr($5,1,50), $NF}'
timeout 900 ncu --set full --import-source on --clock-control none -s 12 -c 9 -o /tmp/step python tools/one_scan.py 64 "because each" k=2 list=1 reps=2 > gpurun_out/r2_ncu_step.log 2>&1
tail -3 gpurun_out/r2_ncu_step.log
bash tools/ncu_export.sh /tmp/step.ncu-rep gpurun_out/r2_step_refine k_refine
cp gpurun_out/r2_step_refine_raw.csv gpurun_out/r2_step_raw.csv
bash tools/ncu_export.sh /tmp/step.ncu-rep gpurun_out/r2_step_front k_front
bash tools/ncu_export.sh /tmp/step.ncu-rep gpurun_out/r2_step_list k_records_list
rm -f gpurun_out/r2_step_front_raw.csv gpurun_out/r2_step_list_raw.csv gpurun_out/r2_step_refine_raw.csv
timeout 600 python tools/path_bench.py > gpurun_out/r2_path_bench.log 2>&1
tail -20 gpurun_out/r2_path_bench.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
cat gpurun_out/r2_bench_ref.json
