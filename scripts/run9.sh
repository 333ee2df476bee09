export LD_LIBRARY_PATH=$PWD/multirobot_pathplanning_benchmark_b200:$LD_LIBRARY_PATH
python scripts/export_blob.py box_stacking /tmp/box_stacking.blob > /dev/null
timeout 200 python -m pytest tests/test_gpu_scene.py -m gpu -x -q 2>&1 | tail -1
for B in 1048576 4194304; do
for v in minb4 base minb4 base; do
  if [ $v = base ]; then L=$PWD/multirobot_pathplanning_benchmark_b200; else L=$PWD/build_variants/$v; fi
  echo "$v $B: $(LD_LIBRARY_PATH=$L timeout 120 native/bench_main /tmp/box_stacking.blob $B 10 | head -1)"
done; done
