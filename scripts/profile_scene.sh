# usage (GPU box): scripts/profile_scene.sh TAG SCENE [B] [KERNEL_REGEX]
# one `ncu --set full` capture of a single launch (after warm-up) of the scene kernel, via native/bench_main
TAG=$1; SCENE=$2; B=${3:-4194304}; K=${4:-check_configs_kernel}
mkdir -p gpurun_out
python scripts/export_blob.py $SCENE /tmp/$SCENE.blob
export LD_LIBRARY_PATH=$PWD/multirobot_pathplanning_benchmark_b200:$LD_LIBRARY_PATH
native/bench_main /tmp/$SCENE.blob $B 5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f \
    -o gpurun_out/prof_${SCENE}_${K}_$TAG native/bench_main /tmp/$SCENE.blob $B 2 > gpurun_out/prof_${SCENE}_${K}_$TAG.log 2>&1
tail -2 gpurun_out/prof_${SCENE}_${K}_$TAG.log
