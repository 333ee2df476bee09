set -x
mkdir -p gpurun_out
: > gpurun_out/r2h_knn_variants.log
for v in tc_b2 tc_b3 tc_b4 tc_b3_p3; do
  echo "== $v" >> gpurun_out/r2h_knn_variants.log
  MRB200_LIB=$PWD/build_variants/$v/libmrb200.so timeout 300 python scripts/knn_check.py 100000 2>&1 | tail -1 >> gpurun_out/r2h_knn_variants.log
  MRB200_LIB=$PWD/build_variants/$v/libmrb200.so timeout 300 python scripts/knn_check.py 100000 euclidean 2>&1 | tail -1 >> gpurun_out/r2h_knn_variants.log
  MRB200_LIB=$PWD/build_variants/$v/libmrb200.so timeout 300 python scripts/knn_check.py 60000 max_euclidean 2 2>&1 | tail -1 >> gpurun_out/r2h_knn_variants.log
done
cut -c1-60,150-260 gpurun_out/r2h_knn_variants.log
timeout 600 python -m pytest tests/test_gpu_knn.py -m gpu -q -x -k "lower_bound or radius" > gpurun_out/r2h_knntests.log 2>&1; tail -3 gpurun_out/r2h_knntests.log
