set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2f_topo.txt 2>&1; head -14 gpurun_out/r2f_topo.txt; nproc; lscpu | grep -i "numa\|socket\|model name" | head
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2f_bench_n$N.json 2> gpurun_out/r2f_bench_n$N.err; tail -c 600 gpurun_out/r2f_bench_n$N.err; python scripts/show_bench.py gpurun_out/r2f_bench_n$N.json 2>/dev/null | head -5; head -c 1500 gpurun_out/r2f_bench_n$N.json
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 8 --steps 2 --warmup 3 --ref-seconds 20 > gpurun_out/r2f_ref_n8.json 2>/dev/null; cut -c1-700 gpurun_out/r2f_ref_n8.json
