set -x
mkdir -p gpurun_out
for N in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2x_bench_n$N.json 2> gpurun_out/r2x_bench_n$N.err; tail -c 300 gpurun_out/r2x_bench_n$N.err; python scripts/show_bench.py gpurun_out/r2x_bench_n$N.json 2>/dev/null | head -3
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 8 --steps 2 --warmup 3 --ref-seconds 20 > gpurun_out/r2x_ref_n8.json 2>/dev/null; cut -c1-200 gpurun_out/r2x_ref_n8.json
