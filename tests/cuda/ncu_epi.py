"""Per-instruction stall listing of the epilogue region of the umma kernel from an ncu report."""
import csv, re, subprocess, collections, sys
rep=sys.argv[1]; cubin=sys.argv[2]; kern=sys.argv[3]; lo=int(sys.argv[4]); hi_=int(sys.argv[5])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if "# Samples" in r][0]
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
lines=[];cur=None;infn=False; sass=[]
for l in dis:
    if l.startswith("//--------------------- .text."):
        infn = kern in l; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur=(m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur); sass.append(l.split("*/",1)[1].strip()[:70])
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
marks=[i for i,l in enumerate(sass) if "USETMAXREG" in l]
epi_start=[i for i in marks if "TRY_ALLOC" in sass[i]][0]
tot=sum(int(data[i][idx["# Samples"]] or 0) for i in range(epi_start,len(sass)))
print("epi start", epi_start, "n", len(sass), "samples", tot)
for i in range(epi_start, len(sass)):
    f,ln=lines[i]
    if f!="gp_umma.cu" and not (lo<=0): pass
    if f=="gp_umma.cu" and not (lo<=ln<=hi_): continue
    if f!="gp_umma.cu": 
        # include inlined helper lines only if neighbours in range
        continue
    s=int(data[i][idx["# Samples"]] or 0)
    st=sorted(((int(data[i][idx[c]] or 0),c[6:]) for c in stall_cols),reverse=True)[:2]
    ex=data[i][idx["Instructions Executed"]]
    print(f"{i:5d} L{ln:<4d} {s:5d} ex={ex:>8} {sass[i]:70s} {[x for x in st if x[0]>0]}")
