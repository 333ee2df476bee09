mkdir -p gpurun_out
L=gpurun_out/r2l_knn.log
: > $L
for v in $VARIANTS; do
  echo "== $v" >> $L
  if [ $v = base ]; then LIB=""; else LIB=$PWD/build_variants/$v/libmrb200.so; fi
  MRB200_LIB=$LIB timeout 300 python scripts/knn_check.py 100000 2>&1 | tail -1 >> $L
  MRB200_LIB=$LIB timeout 300 python scripts/knn_check.py 100000 euclidean 2>&1 | tail -1 >> $L
  MRB200_LIB=$LIB timeout 300 python scripts/knn_check.py 60000 max_euclidean 2 2>&1 | tail -1 >> $L
done
cut -c1-50,78-125,190-300 $L
