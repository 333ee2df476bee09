set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2u_gputests.log 2>&1; tail -4 gpurun_out/r2u_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2u_smoke.log 2>&1; tail -2 gpurun_out/r2u_smoke.log
timeout 1500 python bench.py --verbose > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; tail -c 900 gpurun_out/r2u_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2u_ref.json 2> gpurun_out/r2u_ref.err; cut -c1-300 gpurun_out/r2u_ref.json
