set -x
cd $GRAFT_REPO_ROOT
N=8
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n8.json'))
print(d['n_gpus'], d['ms_per_step'], d['value'], d['matching_records'], d['roofline']['ms_per_launch'], d['roofline']['stage2_ms_per_step'], d['e2e']['value'], d['clocks'])
PY
tail -3 gpurun_out/r2_bench_n$N.err
