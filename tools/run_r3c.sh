mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py tests/test_gpu_device_arith.py tests/test_gpu_overlay.py -m gpu -x -q --timeout 600 2>&1 | tail -5 > gpurun_out/r3c_tests.log; cat gpurun_out/r3c_tests.log
timeout 600 python tools/lsi_variants.py "lsi_fused=0,lsi_cells=0" "lsi_fused=1,lsi_cells=0" "lsi_fused=0,lsi_cells=1" "lsi_fused=1,lsi_cells=1" "lsi_fused=1,lsi_cells=1,lsi_tile_filter=0" > gpurun_out/r3c_variants.jsonl 2> gpurun_out/r3c_variants.err; cat gpurun_out/r3c_variants.jsonl; tail -3 gpurun_out/r3c_variants.err
