mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_lsi_filter_tiles|k_lsi_cells|k_lsi_exact|k_lsi_points" \
  --launch-skip 16 -c 4 -o gpurun_out/r3b_lsi python tools/lsi_variants.py --steps 3 "lsi_tile_filter=1,lsi_cells=1" > gpurun_out/r3b_ncu.log 2>&1
tail -3 gpurun_out/r3b_ncu.log
ncu -i gpurun_out/r3b_lsi.ncu-rep --page raw --csv > gpurun_out/r3b_lsi_raw.csv 2>/dev/null
ncu -i gpurun_out/r3b_lsi.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r3b_lsi_src.csv 2>/dev/null
ls -la gpurun_out/
