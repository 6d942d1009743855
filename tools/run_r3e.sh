mkdir -p gpurun_out; rm -f gpurun_out/r3e_variants.jsonl
for v in a b c d; do
  case $v in a|b) L="6 8 12 16";; c) L="3 4 6";; d) L="12 16 24 32";; esac
  args=""; for n in $L; do args="$args lsi_fused=1,lsi_cells=1,lsi_resolve_ctas=$n"; done
  echo "lib $v" >> gpurun_out/r3e_variants.jsonl
  RJB_LIB=$PWD/rayjoin_b200/librjb200_$v.so timeout 600 python tools/lsi_variants.py $args >> gpurun_out/r3e_variants.jsonl 2> gpurun_out/r3e_variants.err
done
cut -c1-200 gpurun_out/r3e_variants.jsonl
