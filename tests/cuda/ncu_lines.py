"""Joins an ncu report's per-instruction stall samples with nvdisasm line info -> samples per source line.
usage: python tests/cuda/ncu_lines.py report.ncu-rep cubin kernel_substr [topN]"""
import csv, re, subprocess, sys, collections
rep, cubin, kern = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if "# Samples" in r][0]
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
# locate function
lines = []; cur = None; infn = False
for l in dis:
    if l.startswith("//--------------------- .text."):
        infn = kern in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
print("ncu instrs", len(data), "nvdisasm instrs", len(lines))
n = min(len(data), len(lines))
agg = collections.Counter(); execs = collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
st_agg = collections.defaultdict(collections.Counter)
tot = 0
for i in range(n):
    s = int(data[i][idx["# Samples"]] or 0)
    agg[lines[i]] += s; tot += s
    execs[lines[i]] += int(data[i][idx["Instructions Executed"]] or 0)
    for c in stall_cols:
        v = int(data[i][idx[c]] or 0)
        if v: st_agg[lines[i]][c] += v
src = {}
for (f, ln), s in agg.most_common(topn):
    if f not in src:
        try: src[f] = open("/root/repo/acmil_b200/csrc/" + f).read().splitlines()
        except Exception: src[f] = []
    text = src[f][ln - 1].strip()[:90] if ln - 1 < len(src[f]) else ""
    top = ",".join(f"{k[6:]}:{v}" for k, v in st_agg[(f, ln)].most_common(2))
    print(f"{s:6d} {100*s/tot:5.1f}% ex={execs[(f,ln)]:>9} {f}:{ln:<4d} [{top}] {text}")
