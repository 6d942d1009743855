mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device_arith.py tests/test_gpu_reference_pin.py -m gpu -x -q --timeout 600 2>&1 | tail -5 > gpurun_out/r3f_tests.log; cat gpurun_out/r3f_tests.log
timeout 600 python tools/lsi_variants.py "lsi_fused=1,lsi_cells=1" "lsi_fused=1,lsi_cells=0" > gpurun_out/r3f_variants.jsonl 2> gpurun_out/r3f_variants.err; cut -c1-330 gpurun_out/r3f_variants.jsonl; tail -3 gpurun_out/r3f_variants.err
RJB_LIB=$PWD/rayjoin_b200/librjb200_t.so python tools/trace_resolve.py lsi_cells=1 2>&1 | tail -18
