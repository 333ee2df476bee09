mkdir -p gpurun_out
for v in ${TRACE_VARIANTS:-trace_abl1}; do
MRB200_LIB=$PWD/build_variants/$v/libmrb200.so timeout 300 python scripts/knn_one.py 4 > gpurun_out/r2n_$v.log 2>&1
grep TRACE gpurun_out/r2n_$v.log | tail -160 > gpurun_out/r2n_$v.txt
grep "TRACE " gpurun_out/r2n_$v.txt | tail -4
grep "TRACE3" gpurun_out/r2n_$v.txt | tail -48; grep TRACE4 gpurun_out/r2n_$v.txt
done
