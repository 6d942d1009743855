mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device_arith.py -m gpu -x -q --timeout 600 2>&1 | tail -3 > gpurun_out/r3g_tests.log; cat gpurun_out/r3g_tests.log
timeout 600 python tools/lsi_variants.py "lsi_fused=1,lsi_cells=1" "lsi_fused=1,lsi_cells=1,lsi_resolve_ctas=4" "lsi_fused=1,lsi_cells=1,lsi_resolve_ctas=3" "lsi_fused=1,lsi_cells=0" "lsi_fused=0,lsi_cells=1" > gpurun_out/r3g_variants.jsonl 2> gpurun_out/r3g_variants.err; cut -c1-200 gpurun_out/r3g_variants.jsonl; tail -3 gpurun_out/r3g_variants.err
RJB_LIB=$PWD/rayjoin_b200/librjb200_t.so python tools/trace_resolve.py lsi_cells=1 2>&1 | tail -19
