set -x
mkdir -p gpurun_out
N=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2v_bench_n$N.json 2> gpurun_out/r2v_bench_n$N.err; tail -c 500 gpurun_out/r2v_bench_n$N.err; python scripts/show_bench.py gpurun_out/r2v_bench_n$N.json 2>/dev/null | head -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 3 --ref-seconds 20 > gpurun_out/r2v_ref_n$N.json 2>/dev/null; cut -c1-300 gpurun_out/r2v_ref_n$N.json
