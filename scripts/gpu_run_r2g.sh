set -x
mkdir -p gpurun_out
: > gpurun_out/r2g_knn_variants.log
for v in tc_p2 tc_p3 tc_p3_s200 tc_p2_s200 tc_p3_s1000; do
  echo "== $v" >> gpurun_out/r2g_knn_variants.log
  MRB200_LIB=$PWD/build_variants/$v/libmrb200.so timeout 300 python scripts/knn_check.py 100000 2>&1 | tail -1 >> gpurun_out/r2g_knn_variants.log
  MRB200_LIB=$PWD/build_variants/$v/libmrb200.so timeout 300 python scripts/knn_check.py 100000 euclidean 2>&1 | tail -1 >> gpurun_out/r2g_knn_variants.log
  MRB200_LIB=$PWD/build_variants/$v/libmrb200.so timeout 300 python scripts/knn_check.py 60000 max_euclidean 2 2>&1 | tail -1 >> gpurun_out/r2g_knn_variants.log
done
cat gpurun_out/r2g_knn_variants.log
timeout 1200 python -m pytest tests/test_gpu_knn.py -m gpu -q -x > gpurun_out/r2g_knntests.log 2>&1; tail -5 gpurun_out/r2g_knntests.log
