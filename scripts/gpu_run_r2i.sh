set -x
mkdir -p gpurun_out
: > gpurun_out/r2i_knn_variants.log
for v in tc_b1; do
  echo "== $v" >> gpurun_out/r2i_knn_variants.log
  MRB200_LIB=$PWD/build_variants/$v/libmrb200.so timeout 300 python scripts/knn_check.py 100000 2>&1 | tail -1 >> gpurun_out/r2i_knn_variants.log
  MRB200_LIB=$PWD/build_variants/$v/libmrb200.so timeout 300 python scripts/knn_check.py 100000 euclidean 2>&1 | tail -1 >> gpurun_out/r2i_knn_variants.log
  MRB200_LIB=$PWD/build_variants/$v/libmrb200.so timeout 300 python scripts/knn_check.py 60000 max_euclidean 2 2>&1 | tail -1 >> gpurun_out/r2i_knn_variants.log
done
cut -c1-60,150-260 gpurun_out/r2i_knn_variants.log
