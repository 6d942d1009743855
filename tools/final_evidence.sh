#!/bin/bash
# Round evidence on one B200: GPU tests, both bench arms, ncu launch list + full captures.
# Usage: tools/final_evidence.sh TAG      (outputs in gpurun_out/, copied to profiles/ by hand)
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -3 > gpurun_out/${TAG}_gputests.log; cat gpurun_out/${TAG}_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -2 gpurun_out/bench_${TAG}.err
# launch list of the same command (LSI leg only: the PIP leg replays 100 M-point kernels)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench_lbvh.csv \
  python bench.py --legs lsi --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1
# full captures of the query kernels (one launch each, after warm-up launches are skipped)
ncu --set full --clock-control none --import-source on -k regex:"k_lsi_filter|k_lsi_bvh|k_lsi_cells|k_lsi_resolve|k_lsi_exact|k_lsi_points" \
  --launch-skip 6 -c 3 -o gpurun_out/${TAG}_lsi_full python bench.py --legs lsi --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${TAG}_ncu_lsi.log 2>&1
ncu -i gpurun_out/${TAG}_lsi_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_lsi_kernels_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:"k_pip_grid|k_pip_bvh|k_sort_onesweep_packed|k_grid_lsi" \
  -c 6 -o gpurun_out/${TAG}_pip_full python tools/pip_bench.py --points 50000000 --modes grid,lbvh --sort 1 --park 1 --check 0 --repeat 1 --grid-size 16384 > gpurun_out/${TAG}_ncu_pip.log 2>&1
ncu -i gpurun_out/${TAG}_pip_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_pip_kernels_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_lsi_full.ncu-rep gpurun_out/${TAG}_pip_full.ncu-rep
python tools/pip_bench.py --points 100000000 --modes grid,lbvh --sort 1 --park 1 --check -1 --grid-size 16384 2>/dev/null > gpurun_out/pip_${TAG}.jsonl
cat gpurun_out/pip_${TAG}.jsonl | cut -c1-400
