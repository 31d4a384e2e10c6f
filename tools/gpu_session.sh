set -x
cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2r_tests.log
tail -8 gpurun_out/r2r_tests.log
timeout 1200 python bench.py > gpurun_out/r2r_bench_n1.json 2> gpurun_out/r2r_bench_n1.err
cat gpurun_out/r2r_bench_n1.json; tail -5 gpurun_out/r2r_bench_n1.err
