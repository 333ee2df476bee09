mkdir -p gpurun_out
cap() {  # name lib
    MRB200_LIB=$2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -s 2 -c 1 -f \
        -o gpurun_out/cap_r2m_$1 python scripts/prof_driver.py knn 100000 tensor > gpurun_out/cap_r2m_$1.log 2>&1
    tail -1 gpurun_out/cap_r2m_$1.log
}
cap knn_tc ""
cap knn_tc_abl1 $PWD/build_variants/abl1/libmrb200.so
