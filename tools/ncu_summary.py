"""Summarise an ncu report: key raw metrics + top CUDA source lines + opcode mix (development aid)."""
import csv, subprocess, sys, re, collections, io, json, os
rep = sys.argv[1]
# optional: --json <file> [--meta key=value ...]: machine-readable sidecar bench.py reads (profiles/*.json)
json_out = None; meta = {}
if "--json" in sys.argv:
    i = sys.argv.index("--json"); json_out = sys.argv[i + 1]
    for kv in sys.argv[i + 2:]:
        if "=" in kv:
            k_, v_ = kv.split("=", 1); meta[k_] = v_
    del sys.argv[i:]
side = {"report": os.path.basename(rep), "meta": meta, "metrics": {}, "opcodes": {}}
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct', 'sm__inst_executed_pipe_fma.avg.pct',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'smsp__issue_active.avg.pct',
        'smsp__inst_executed.sum ', 'smsp__thread_inst_executed_per_inst_executed', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'smsp__average_warps_issue_stalled', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct', 'launch__waves']
for r in rows[2:]:
    print("==", r[hdr.index('Kernel Name')][:90] if 'Kernel Name' in hdr else '')
    for h, u, v in zip(hdr, units, r):
        if any(h.startswith(k.strip()) if k.endswith(' ') else (k in h) for k in keys):
            try: side["metrics"][h + " [" + u + "]"] = float(v.replace(",", ""))
            except Exception: pass
            if 'stalled' in h and float(v or 0) < 0.15: continue
            print(f"  {h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
agg = []; fname = None; op = collections.Counter(); tot_sass = 0
for r in rows:
    if r and r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if len(r) < 8: continue
    if r[0].isdigit():
        try: agg.append((fname, int(r[0]), r[1].strip()[:100], float(r[7]), float(r[6])))
        except Exception: pass
    elif r[0] == '' and r[2].startswith('0x'):
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[3])
        try:
            n = float(r[7]); op[m.group(2) if m else '?'] += n; tot_sass += n
        except Exception: pass
tot = sum(a[3] for a in agg) or 1; tots = sum(a[4] for a in agg) or 1
print("-- top source lines (instr %, stall-sample %)")
for a in sorted(agg, key=lambda a: -a[3])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{a[3]/tot*100:5.1f}% {a[4]/tots*100:5.1f}%  {a[0]}:{a[1]}: {a[2]}")
side["opcodes"] = {o: n for o, n in op.most_common(40)}   # warp-level executed instruction counts by opcode
side["warp_instructions"] = tot_sass
if json_out:
    with open(json_out, "w") as f: json.dump(side, f, indent=1)
print("-- opcode mix")
print("  ".join(f"{o}:{n/tot_sass*100:.1f}%" for o, n in op.most_common(16)))
