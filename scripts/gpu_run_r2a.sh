set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
python __graft_entry__.py --smoke > gpurun_out/r2a_smoke.log 2>&1; tail -2 gpurun_out/r2a_smoke.log
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_gputests.log 2>&1; tail -15 gpurun_out/r2a_gputests.log
timeout 1500 python bench.py --verbose > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.err; wc -c gpurun_out/r2a_bench.json
scripts/capture_all.sh r2a
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-extra --no-cpu > gpurun_out/r2a_launches_bench.log 2>&1
