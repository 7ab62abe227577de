"""Per-source-line view of an ncu report (development aid): samples %, instructions %, cycles per instruction, top stall reasons.
   python tools/ncu_lines.py report.ncu-rep [min_pct]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
fname = None; hdr = None; out = []
for r in rows:
    if r and r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if not r or not r[0].isdigit() or hdr is None: continue
    d = dict(zip(hdr, r))
    try:
        smp = float(d['# Samples'] or 0); ins = float(d['Instructions Executed'] or 0)
    except Exception:
        continue
    if smp == 0 and ins == 0: continue
    st = {k[6:]: float(d[k] or 0) for k in hdr if k.startswith('stall_') and 'Not Issued' not in k}
    out.append((fname, int(r[0]), r[1].strip()[:70], smp, ins, st))
ts = sum(o[3] for o in out); ti = sum(o[4] for o in out)
print(f"total samples {ts:.0f}, warp instructions {ti:.0f}")
tot_st = collections.Counter()
for o in out:
    for k, v in o[5].items(): tot_st[k] += v
print("stall totals:", "  ".join(f"{k}:{v/ts*100:.1f}%" for k, v in tot_st.most_common(10)))
cum = 0
for o in out:
    cum += o[3]
    if o[3] / ts * 100 < minp and o[4] / ti * 100 < minp: continue
    top = sorted(o[5].items(), key=lambda kv: -kv[1])[:3]
    print(f"{o[0][:14]:14s}:{o[1]:4d} smp {o[3]/ts*100:5.2f}% (cum {cum/ts*100:5.1f}) ins {o[4]/ti*100:5.2f}%  " +
          " ".join(f"{k}:{v/max(o[3],1)*100:.0f}" for k, v in top) + f"  | {o[2]}")
