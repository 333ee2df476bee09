mkdir -p gpurun_out
echo "== latency kernel on" > gpurun_out/r2t_single.log
timeout 300 python scripts/bench_single_query.py >> gpurun_out/r2t_single.log 2>&1
echo "== latency kernel off" >> gpurun_out/r2t_single.log
MRB200_NO_LATENCY_KERNEL=1 timeout 300 python scripts/bench_single_query.py >> gpurun_out/r2t_single.log 2>&1
cat gpurun_out/r2t_single.log
timeout 900 python -m pytest tests/test_gpu_scene.py tests/test_gpu_smoke.py tests/test_gpu_reference_planners.py -x -q -m gpu 2>&1 | tail -3
