#!/bin/bash
# compute-sanitizer over a slice of the GPU test-suite (small datasets: the tools slow kernels
# down 10-100x): memcheck and racecheck summaries into gpurun_out/.
TAG=${1:-r2}
SEL='tests/test_gpu_parity.py tests/test_gpu_sort.py tests/test_gpu_overlay.py -k "(voronoi or lattice or shared or tiny or packed or sort_pairs) and not dense and not 3_000_001 and not 5_000_001 and not binary"'
for tool in memcheck racecheck; do
  eval timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 3 \
    python -m pytest $SEL -q -x --timeout 1400 > gpurun_out/${TAG}_sanitizer_${tool}.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/${TAG}_sanitizer_${tool}.log | tail -3
done
