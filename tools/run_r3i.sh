mkdir -p gpurun_out
RJB_LIB=$PWD/rayjoin_b200/librjb200_a.so timeout 600 python tools/lsi_variants.py "lsi_fused=1,lsi_cells=1" "lsi_fused=1,lsi_cells=1,lsi_resolve_ctas=4" 2>/dev/null | cut -c1-200
RJB_LIB=$PWD/rayjoin_b200/librjb200_b.so python tools/trace_resolve.py lsi_cells=1 2>&1 | grep -E "CTA lifetime|final|tail: barrier|kernel span|mark  [2345] "
