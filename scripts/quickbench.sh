# usage: scripts/quickbench.sh TAG   (on the GPU box) -- scene parity tests, then a short bench without the CPU leg
TAG=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scene.py -m gpu -q 2>&1 | tail -4
timeout 900 python bench.py --no-cpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -3 gpurun_out/bench_$TAG.err
python scripts/show_bench.py gpurun_out/bench_$TAG.json
