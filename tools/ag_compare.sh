#!/bin/bash
# Adaptive leaf grouping against fixed leaves on the headline workload and on a shared-chain variant
for sc in 0 0.05; do
  for cfg in "--ag 0 --leaf-size 4" "--ag 0 --leaf-size 8" "--ag 1 --ag-iter 5 --enlarge 3.5" "--ag 1 --ag-iter 5 --enlarge 5" "--ag 1 --ag-iter 2 --enlarge 3.5"; do
    python bench.py --legs lsi --steps 20 --warmup 5 --no-cpu-baseline --share-chains $sc $cfg 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(json.dumps({'share_chains': $sc, 'cfg': '$cfg', 'leaves': d['index_units'], 'index_MB': round(d['index_bytes']/1e6,1), 'build_ms': round(d['index_build_ms'],3), 'query_ms': round(d['ms_per_step'],4), 'candidates': d['candidate_pairs'], 'pairs': d['result_pairs'], 'kernel_ms': {k: round(v,4) for k,v in d['kernel_ms'].items()}}))"
  done
done
