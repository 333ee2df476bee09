set -x
mkdir -p gpurun_out
timeout 300 python scripts/knn_check.py 100000 > gpurun_out/r2c_knn_check.log 2>&1; tail -2 gpurun_out/r2c_knn_check.log
timeout 300 python scripts/knn_check.py 100000 euclidean >> gpurun_out/r2c_knn_check.log 2>&1; tail -2 gpurun_out/r2c_knn_check.log
timeout 300 python scripts/knn_check.py 60000 max_euclidean 2 >> gpurun_out/r2c_knn_check.log 2>&1; tail -2 gpurun_out/r2c_knn_check.log
timeout 300 python scripts/knn_check.py 50000 max_euclidean 3 >> gpurun_out/r2c_knn_check.log 2>&1; tail -2 gpurun_out/r2c_knn_check.log
timeout 1200 python -m pytest tests/test_gpu_knn.py -m gpu -q > gpurun_out/r2c_knntests.log 2>&1; tail -5 gpurun_out/r2c_knntests.log
TAG=r2c
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -s 2 -c 1 -f -o gpurun_out/cap_${TAG}_knn_tc python scripts/prof_driver.py knn 100000 tensor > gpurun_out/cap_${TAG}_knn.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2c_launches_knn.csv python scripts/prof_driver.py knn 100000 tensor > /dev/null 2>&1
