# usage: scripts/build_variant.sh NAME "-DFOO=1 ..." [source file stem, default scene_kernels]
#   -> build_variants/NAME/libmrb200.so (experiment builds, git-ignored)
set -e
NAME=$1; DEFS=$2; WHICH=${3:-scene_kernels}
mkdir -p build_variants/$NAME
cd multirobot_pathplanning_benchmark_b200/csrc
OBJS=""
for f in capi scene_kernels abstract_kernels knn_kernels knn_tc_kernels; do
  if [ $f = $WHICH ] || [ ! -f ../build/$f.cu.o ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $DEFS -c $f.cu -o ../../build_variants/$NAME/$f.o
    OBJS="$OBJS ../../build_variants/$NAME/$f.o"
  else
    OBJS="$OBJS ../build/$f.cu.o"
  fi
done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../build_variants/$NAME/libmrb200.so $OBJS
echo built build_variants/$NAME
