"""Summarise an .ncu-rep: headline metrics, stall reasons, per-function and per-line shares."""
import csv, re, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 22
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for k in keys:
    for i, h in enumerate(hdr):
        if h == k:
            print(f"{k:62s} {vals[i]} {units[i]}")
st = [(float(vals[i].replace(',', '')), h) for i, h in enumerate(hdr)
      if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
print("stall reasons (warps per issue-active):", ", ".join(
    f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}={v:.2f}"
    for v, h in sorted(st, reverse=True)[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur = None
agg = []
for r in csv.reader(src.splitlines()):
    if not r:
        continue
    if r[0] == 'File Path':
        cur = r[1]
        continue
    if r[0] in ('Function Name', 'Line No'):
        continue
    if r[0].isdigit():
        try:
            agg.append((int(r[4].replace(',', '')), int(r[7].replace(',', '')), cur, int(r[0]), r[1].strip()[:96]))
        except Exception:
            pass
tots, toti = sum(a[0] for a in agg) or 1, sum(a[1] for a in agg) or 1
print(f"samples {tots} instructions {toti}")
files = {}
for s_, i_, f, l, _ in agg:
    files.setdefault(f, []).append((l, s_, i_))
reg = {}
for f, items in files.items():
    try:
        lines = open(f).read().split('\n')
    except Exception:
        lines = []
    marks = [(i + 1, l.strip()[:64]) for i, l in enumerate(lines)
             if re.match(r'\s*(static |template.*|__global__ |__device__ )?(CORAL_(DEV|HD)|__global__|__device__) ', l)]
    marks.append((10 ** 9, 'END'))
    for l, s_, i_ in items:
        name = f.split('/')[-1]
        for (a, n), (b, _) in zip(marks, marks[1:]):
            if a <= l < b:
                name = n
                break
        x = reg.setdefault(name, [0, 0])
        x[0] += s_
        x[1] += i_
print("--- per function: inst% samples%")
for k, (s_, i_) in sorted(reg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{100 * i_ / toti:5.1f} {100 * s_ / tots:5.1f}  {k}")
print("--- top lines by stall samples: samples% inst%")
for a in sorted(agg, reverse=True)[:top]:
    print(f"{100 * a[0] / tots:5.1f} {100 * a[1] / toti:5.1f}  {a[2].split('/')[-1]}:{a[3]}  {a[4]}")
