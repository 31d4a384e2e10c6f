set -x
cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2k_tests.log
tail -8 gpurun_out/r2k_tests.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_front_exact -s 1 -c 1 -f -o /tmp/r2k_exact python tools/one_scan.py 8 "the" > gpurun_out/r2k_ncu_exact.log 2>&1
tools/ncu_export.sh /tmp/r2k_exact.ncu-rep gpurun_out/r2k_exact
