#!/bin/bash
# GPU box: `ncu --set full` captures of the kernels the bench line quotes, on the CURRENT build, plus the launch list of
# the bench command.  usage: scripts/capture_all.sh TAG   ->  gpurun_out/cap_TAG_*.ncu-rep (+ .log)
TAG=${1:-r2}
mkdir -p gpurun_out
cap() {  # name kernel-regex skip -- command...
    local name=$1 k=$2 skip=$3; shift 3
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f \
        -o gpurun_out/cap_${TAG}_$name "$@" > gpurun_out/cap_${TAG}_$name.log 2>&1
    tail -1 gpurun_out/cap_${TAG}_$name.log
}
cap configs_box_rearrangement_4M check_configs_kernel 4 python scripts/prof_driver.py configs box_rearrangement 4194304
cap configs_box_stacking_1M check_configs_kernel 4 python scripts/prof_driver.py configs box_stacking 1048576
cap configs_mobile_wall_four_2M check_configs_kernel 4 python scripts/prof_driver.py configs mobile_wall_four 2097152
cap edges_box_rearrangement_local check_edges_kernel 3 python scripts/prof_driver.py edges box_rearrangement 131072 local
cap edges_box_rearrangement_uniform check_edges_kernel 3 python scripts/prof_driver.py edges box_rearrangement 16384 uniform
cap edges_box_stacking_local check_edges_kernel 3 python scripts/prof_driver.py edges box_stacking 65536 local
cap knn_tc knn_tc_kernel 2 python scripts/prof_driver.py knn 100000 tensor
cap radius_tc knn_tc_kernel 2 python scripts/prof_driver.py radius 100000
cap radius_filter radius_tc_filter_warp_kernel 1 python scripts/prof_driver.py radius 100000
cap knn_rerank knn_rerank_kernel 2 python scripts/prof_driver.py knn 100000 tensor
# summaries on the box (gpurun brings back at most 64 MiB): profiles/TAG_*.txt + traffic.json, then drop the large reports
# except the two whose source pages are read here
python scripts/ncu_to_traffic.py $TAG > gpurun_out/${TAG}_ncu_to_traffic.log 2>&1
mkdir -p gpurun_out/profiles_$TAG
cp profiles/${TAG}_*.txt profiles/traffic.json gpurun_out/profiles_$TAG/
for f in gpurun_out/cap_${TAG}_*.ncu-rep; do
  case $f in *knn_tc.ncu-rep|*configs_box_rearrangement_4M.ncu-rep) ;; *) rm -f $f ;; esac
done
# launch list of the default bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-planners > gpurun_out/${TAG}_launches_bench.log 2>&1
tail -1 gpurun_out/${TAG}_launches_bench.log | cut -c1-300
