mkdir -p gpurun_out; rm -f gpurun_out/r3h_variants.jsonl
for v in a c; do
  echo "lib $v" >> gpurun_out/r3h_variants.jsonl
  RJB_LIB=$PWD/rayjoin_b200/librjb200_$v.so timeout 600 python tools/lsi_variants.py "lsi_fused=1,lsi_cells=1" >> gpurun_out/r3h_variants.jsonl 2> gpurun_out/r3h_variants.err
done
cut -c1-200 gpurun_out/r3h_variants.jsonl
RJB_LIB=$PWD/rayjoin_b200/librjb200_b.so python tools/trace_resolve.py lsi_cells=1 2>&1 | tail -19
