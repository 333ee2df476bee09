mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2s_launches_knn.csv python scripts/prof_driver.py knn 100000 tensor > gpurun_out/r2s_knn.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2s_launches_knn.csv")) if len(r)>5 and r[0].isdigit()]
for r in rows[-7:]:
    print("  ", r[4][:70], r[-1])
PY
VARIANTS="base" timeout 600 bash scripts/gpu_run_r2l.sh > /dev/null; grep -o "== .*\|uncertified_rows.: [0-9]*\|tensor [0-9.]* ms" gpurun_out/r2l_knn.log | paste - - - - - - -
