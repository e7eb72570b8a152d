"""Top stall-sample source lines of an ncu report (which lines the warps wait on, and why)."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = None; cur = None; per = {}
reasons = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_mio", "stall_lg", "stall_not_selected",
           "stall_branch_resolving", "stall_dispatch", "stall_no_inst", "stall_selected"]
tot_r = {r: 0 for r in reasons}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No" and len(r) > 5: hdr = r; idx = {k: hdr.index(k) for k in reasons + ["Warp Stall Sampling (All Samples)"]}; continue
    if hdr and len(r) == len(hdr) and r[0]:
        try: s = int(r[idx["Warp Stall Sampling (All Samples)"]])
        except ValueError: continue
        e = per.setdefault((cur, int(r[0])), [0, r[1], {k: 0 for k in reasons}])
        e[0] += s
        for k in reasons:
            try: v = int(r[idx[k]])
            except ValueError: v = 0
            e[2][k] += v; tot_r[k] += v
tot = sum(v[0] for v in per.values())
print("total samples", tot, {k: round(v / max(tot, 1) * 100, 1) for k, v in tot_r.items() if v})
for (f, l), (n, s, rs) in sorted(per.items(), key=lambda x: -x[1][0])[:top]:
    main = sorted(rs.items(), key=lambda x: -x[1])[:2]
    print(f"{n / tot * 100:5.1f}%  {f}:{l}: {s[:70]:70s} {main[0][0][6:]}={main[0][1]} {main[1][0][6:]}={main[1][1]}")
