set -x
cd $GRAFT_REPO_ROOT
timeout 400 python bench.py > gpurun_out/r2final_bench_n1.json 2> gpurun_out/r2final_bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2final_bench_n1.json') if l.startswith('{')][0])
print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline_step']['frac'], d['e2e']['value'], d['e2e'].get('fd',{}).get('value'), d['clocks'])
for s in d['secondary_workloads']: print(s['config'][:50], s['ms'], s['roofline']['frac'])
PY
tail -2 gpurun_out/r2final_bench_n1.err
