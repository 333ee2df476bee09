# usage (GPU box): scripts/ab_two_phase.sh  -- scene parity tests, then native/bench_main with the two-phase tiles forced off / on / adaptive
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scene.py tests/test_gpu_scene_fullsize.py tests/test_gpu_smoke.py -m gpu -x -q 2>&1 | tail -5
MRB200_TWO_PHASE=1 timeout 600 python -m pytest tests/test_gpu_scene.py tests/test_gpu_scene_fullsize.py -m gpu -x -q 2>&1 | tail -2
export LD_LIBRARY_PATH=$PWD/multirobot_pathplanning_benchmark_b200:$LD_LIBRARY_PATH
for s in box_rearrangement mobile_wall_four box_stacking 2d_handover; do python scripts/export_blob.py $s /tmp/$s.blob > /dev/null; done
for tp in 0 1 auto; do
  for s in box_rearrangement mobile_wall_four box_stacking; do
    if [ $tp = auto ]; then unset MRB200_TWO_PHASE; else export MRB200_TWO_PHASE=$tp; fi
    echo "two_phase=$tp $s: $(timeout 120 native/bench_main /tmp/$s.blob 4194304 10 | tr '\n' ' ')"
  done
done
