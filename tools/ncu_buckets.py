"""Summarise an ncu report: key raw metrics + executed instructions per source line.

    python tools/ncu_buckets.py REPORT.ncu-rep N_PARTICLES [TOP_LINES] [MODEL]

With MODEL (rw1d / lg42 / lg63) the DRAM traffic of the captured launch is recorded in
profiles/traffic.json, which bench.py reports as roofline.traffic when it runs at the same N."""
import csv, json, os, subprocess, sys
rep, nparticles = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); h = rows[0]; r = rows[2]
for k in ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__inst_executed.avg.per_cycle_elapsed',
          'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
          'launch__registers_per_thread', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
          'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
          'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
          'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']:
    if k in h: print(f"{k:95s} {r[h.index(k)]} {rows[1][h.index(k)]}")
if len(sys.argv) > 4:
    def tobytes(k):
        v, u = float(r[h.index(k)]), rows[1][h.index(k)].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    tp = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "traffic.json")
    try: tj = json.load(open(tp))
    except Exception: tj = {}
    tj[sys.argv[4]] = {"particles": int(nparticles), "dram_bytes": tobytes('dram__bytes_read.sum') + tobytes('dram__bytes_write.sum'),
                       "dram_bytes_read": tobytes('dram__bytes_read.sum'), "dram_bytes_write": tobytes('dram__bytes_write.sum'),
                       "kernel_us_under_ncu": float(r[h.index('gpu__time_duration.sum')]), "report": os.path.basename(rep)}
    json.dump(tj, open(tp, "w"), indent=1, sort_keys=True)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur = None; per = {}; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No" and len(r) > 5: hdr = r; ii = hdr.index("Instructions Executed"); continue
    if hdr and len(r) > ii and r[0] != "":
        try: n = int(r[ii])
        except ValueError: continue
        e = per.setdefault((cur, int(r[0])), [0, r[1]]); e[0] += n
tot = sum(v[0] for v in per.values())
print("total lane-instr/particle", tot * 32 / nparticles)
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
for (f, l), (n, s) in sorted(per.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{n / tot * 100:5.1f}% {n * 32 / nparticles:6.1f}  {f}:{l}: {s[:100]}")

# ---- buckets by function (line ranges of the current sources)
import re
def func_ranges(path):
    """crude: map line -> enclosing function name by scanning for '__device__'/'__global__' definitions"""
    out = []; name = None
    for i, line in enumerate(open(path), 1):
        m = re.search(r'(?:__device__|__global__)[^;{]*?\b(\w+)\s*\(', line)
        if m and not line.strip().startswith("//"): name = m.group(1)
        m2 = re.match(r'\s*(?:template <[^>]*>\s*)?(?:struct)\s+(\w+)', line)
        if m2: name = m2.group(1)
        out.append(name)
    return out
import os
base = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "probzelus-haskell_b200", "csrc")
fr = {f: func_ranges(os.path.join(base, f)) for f in ("pz_kernels.cu", "pz_math.cuh")}
b = {}
for (f, l), (n, s) in per.items():
    key = fr[f][l - 1] if f in fr and l - 1 < len(fr[f]) else f
    b[key] = b.get(key, 0) + n
print("---- by function")
for k, v in sorted(b.items(), key=lambda x: -x[1]):
    print(f"{v / tot * 100:5.1f}% {v * 32 / nparticles:6.1f}  {k}")
