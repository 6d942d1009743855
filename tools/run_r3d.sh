mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_lsi_resolve" \
  --launch-skip 6 -c 1 -o gpurun_out/r3d_res python tools/lsi_variants.py --steps 3 "lsi_fused=1,lsi_cells=1" > gpurun_out/r3d_ncu.log 2>&1
tail -2 gpurun_out/r3d_ncu.log
ncu -i gpurun_out/r3d_res.ncu-rep --page raw --csv > gpurun_out/r3d_res_raw.csv 2>/dev/null
ncu -i gpurun_out/r3d_res.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r3d_res_src.csv 2>/dev/null
rm -f gpurun_out/r3d_res.ncu-rep
