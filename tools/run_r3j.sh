mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -4 > gpurun_out/r3j_tests.log; cat gpurun_out/r3j_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r3j.json 2> gpurun_out/bench_r3j.err; tail -2 gpurun_out/bench_r3j.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_r3j.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','kernel_ms','index_build_ms','index_bytes','parity_vs_oracle','gpu_launches')})
print(d['e2e']); print(d['roofline']); print(d['roofline_query'])
print([ (r['kernel'], round(r['ms'],4), round(r['frac'],3)) for r in d['roofline_kernels']])
print(d.get('pip',{}).get('query_ms'), d.get('overlay',{}).get('lbvh',{}).get('device_total_ms'), d.get('overlay',{}).get('grid',{}).get('device_total_ms'))
P
