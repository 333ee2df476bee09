# usage (GPU box): scripts/run_variants.sh NAME[@WARPS]...   -- native/bench_main on three scenes for each variant library
for s in box_rearrangement mobile_wall_four box_stacking; do python scripts/export_blob.py $s /tmp/$s.blob > /dev/null; done
for vw in "$@"; do
  v=${vw%@*}; w=""; [ "$v" != "$vw" ] && w=${vw#*@}
  for s in box_rearrangement mobile_wall_four box_stacking; do
    echo "$vw $s: $(MRB200_WARPS=$w LD_LIBRARY_PATH=$PWD/build_variants/$v native/bench_main /tmp/$s.blob 2097152 10 | tr '\n' ' ')"
  done
done
