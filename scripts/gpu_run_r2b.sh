set -x
mkdir -p gpurun_out
timeout 300 python scripts/knn_check.py 20000 > gpurun_out/r2b_knn_check.log 2>&1; tail -3 gpurun_out/r2b_knn_check.log
timeout 300 python scripts/knn_check.py 100000 >> gpurun_out/r2b_knn_check.log 2>&1; tail -2 gpurun_out/r2b_knn_check.log
timeout 300 python scripts/knn_check.py 100000 euclidean >> gpurun_out/r2b_knn_check.log 2>&1; tail -2 gpurun_out/r2b_knn_check.log
timeout 300 python scripts/knn_check.py 60000 max_euclidean 2 >> gpurun_out/r2b_knn_check.log 2>&1; tail -2 gpurun_out/r2b_knn_check.log
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2b_gputests.log 2>&1; tail -25 gpurun_out/r2b_gputests.log
for s in 1 2 3; do timeout 300 python scripts/ref_planner_run.py box_stacking composite_prm --device cuda --seeds $s --max-time 120 2>&1 | tail -1 | cut -c1-900; done > gpurun_out/r2b_prm.log; cat gpurun_out/r2b_prm.log
TAG=r2b
timeout 600 ncu --set full --clock-control none --import-source on -k regex:check_configs_kernel -s 4 -c 1 -f -o gpurun_out/cap_${TAG}_configs_box_rearrangement_4M python scripts/prof_driver.py configs box_rearrangement 4194304 > gpurun_out/cap_${TAG}_configs.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -s 2 -c 1 -f -o gpurun_out/cap_${TAG}_knn_tc python scripts/prof_driver.py knn 100000 tensor > gpurun_out/cap_${TAG}_knn.log 2>&1
