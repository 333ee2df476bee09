set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2o_gputests.log 2>&1; tail -5 gpurun_out/r2o_gputests.log
timeout 1500 python bench.py --verbose > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; tail -c 600 gpurun_out/r2o_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.log 2>&1; tail -2 gpurun_out/r2o_smoke.log
