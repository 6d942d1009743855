mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py tests/test_gpu_device_arith.py -m gpu -x -q --timeout 600 2>&1 | tail -5 > gpurun_out/r3a_tests.log; cat gpurun_out/r3a_tests.log
timeout 600 python tools/lsi_variants.py "lsi_tile_filter=0,lsi_cells=0" "lsi_tile_filter=1,lsi_cells=0" "lsi_tile_filter=0,lsi_cells=1" "lsi_tile_filter=1,lsi_cells=1" > gpurun_out/r3a_variants.jsonl 2> gpurun_out/r3a_variants.err; cat gpurun_out/r3a_variants.jsonl; tail -3 gpurun_out/r3a_variants.err
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
