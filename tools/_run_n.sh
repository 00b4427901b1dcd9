ncu --metrics gpu__time_duration.sum --clock-control none -k regex:beam_search -c 12 --csv --log-file gpurun_out/r2_n_launches.csv python tools/beam_perf.py --utts 8192 --iters 3 > gpurun_out/r2_n.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2_n_launches.csv")) if len(r)>5 and r[0].isdigit()]
for r in rows: print(r[4][:70], r[-1])
PY
python tools/beam_perf.py --utts 8192 --iters 5 2>&1 | grep utts
