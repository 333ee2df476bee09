mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2q_launches_radius.csv python scripts/prof_driver.py radius 100000 > gpurun_out/r2q_radius.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2q_launches_knn.csv python scripts/prof_driver.py knn 100000 tensor > gpurun_out/r2q_knn.log 2>&1
python - <<'PY'
import csv
for f in ("radius","knn"):
    rows=[r for r in csv.reader(open(f"gpurun_out/r2q_launches_{f}.csv")) if len(r)>5 and r[0].isdigit()]
    print(f)
    for r in rows[-14:]:
        print("  ", r[4][:70], r[-1], r[-2])
PY
