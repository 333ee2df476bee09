set -x
mkdir -p gpurun_out
L=gpurun_out/r2j_knn_ablation.log
: > $L
for v in base abl1 abl2; do
  echo "== $v" >> $L
  if [ $v = base ]; then LIB=""; else LIB=$PWD/build_variants/$v/libmrb200.so; fi
  MRB200_LIB=$LIB timeout 300 python scripts/knn_check.py 100000 2>&1 | tail -1 >> $L
  MRB200_LIB=$LIB timeout 300 python scripts/knn_check.py 100000 euclidean 2>&1 | tail -1 >> $L
done
cut -c1-70,150-260 $L
