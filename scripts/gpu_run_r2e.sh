set -x
mkdir -p gpurun_out
timeout 300 python scripts/knn_check.py 100000 > gpurun_out/r2e_knn_check.log 2>&1; tail -2 gpurun_out/r2e_knn_check.log
timeout 1200 python -m pytest tests/test_gpu_knn.py -m gpu -q -x > gpurun_out/r2e_knntests.log 2>&1; tail -8 gpurun_out/r2e_knntests.log
timeout 900 bash scripts/run_variants.sh base pad4 pad4_m10 m10 m12 pad4_efk efk > gpurun_out/r2e_variants.log 2>&1; cat gpurun_out/r2e_variants.log
timeout 600 python -m pytest tests/test_gpu_scene.py -m gpu -q -x > gpurun_out/r2e_scenetests.log 2>&1; tail -3 gpurun_out/r2e_scenetests.log
timeout 1500 python bench.py --verbose > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -c 1500 gpurun_out/r2e_bench.err
