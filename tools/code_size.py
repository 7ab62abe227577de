"""Static instruction count of a kernel by source region, from nvdisasm line info (development aid):
   python tools/code_size.py <object or cubin> <kernel-name-substring>"""
import re, collections, subprocess, sys, os, tempfile
obj, pat = sys.argv[1], sys.argv[2]
with tempfile.TemporaryDirectory() as d:
    if not obj.endswith(".cubin"):
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
        obj = os.path.join(d, [f for f in os.listdir(d) if f.endswith(".cubin")][0])
    dis = subprocess.run(["nvdisasm", "-g", obj], capture_output=True, text=True).stdout
cur = None; stat = collections.Counter(); on = False
for line in dis.splitlines():
    if line.startswith(".text."):
        on = pat in line
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', line):
        stat[cur] += 1
tot = sum(stat.values()); print("total instructions", tot, f"= {tot*16/1024:.1f} KB")
def phase(k):
    if k is None: return "none"
    f, l = k
    return f if f != 'tpp_kernel.cuh' else f"tpp:{l//50*50:04d}"
ph = collections.Counter()
for k, v in stat.items(): ph[phase(k)] += v
for k, v in sorted(ph.items(), key=lambda kv: -kv[1])[:24]: print(f"{k:30s} {v:6d} {v*16/1024:6.1f} KB")
print("top lines")
for k, v in stat.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 24): print(v, k)
