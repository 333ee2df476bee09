set -x
mkdir -p gpurun_out
L=gpurun_out/r2k_knn.log
: > $L
for v in base fast0; do
  echo "== $v" >> $L
  if [ $v = base ]; then LIB=""; else LIB=$PWD/build_variants/$v/libmrb200.so; fi
  MRB200_LIB=$LIB timeout 300 python scripts/knn_check.py 100000 2>&1 | tail -2 >> $L
  MRB200_LIB=$LIB timeout 300 python scripts/knn_check.py 100000 euclidean 2>&1 | tail -2 >> $L
  MRB200_LIB=$LIB timeout 300 python scripts/knn_check.py 60000 max_euclidean 2 2>&1 | tail -2 >> $L
  MRB200_LIB=$LIB timeout 300 python scripts/knn_check.py 50000 max_euclidean 3 2>&1 | tail -2 >> $L
done
timeout 900 python -m pytest tests/test_gpu_knn.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2k_knntests.log
cut -c1-70,140-300 $L
tail -5 gpurun_out/r2k_knntests.log
