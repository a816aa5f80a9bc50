"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`: warp-instructions executed and stall
samples per CUDA source line (summed over the SASS attributed to it).  usage: ncu_source_summary.py file.csv [top]"""
import csv, sys, collections
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
hdr = None
agg = {}
tot_i = tot_s = 0
stalls_total = collections.Counter()
opc = collections.Counter()
cur_line = (None, -1, "")
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    if r[0].isdigit():   # a CUDA source line (its own metric columns are aggregates and may be shifted by commas: ignored)
        cur_line = (cur_file, int(r[0]), r[1].strip()[:100]); continue
    if r[0] != "" or len(r) != len(hdr) or not r[2].startswith("0x"): continue
    sass = r[3]
    d = {h: v for h, v in list(zip(hdr, r))[4:]}
    try:
        inst = int(d.get("Instructions Executed") or 0); samp = int(d.get("# Samples") or 0)
    except ValueError:
        continue
    a = agg.setdefault(cur_line, [0, 0, collections.Counter()])
    a[0] += inst; a[1] += samp
    tot_i += inst; tot_s += samp
    t = sass.strip().split()
    opc[(t[1] if t[0].startswith("@") else t[0]).rstrip(";")] += inst
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0", "-"):
            a[2][k] += int(v); stalls_total[k] += int(v)
print(f"total warp-instructions {tot_i}, samples {tot_s}")
print("stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in stalls_total.most_common(12)))
print("top opcodes:", ", ".join(f"{k}={100*v/tot_i:.1f}%" for k, v in opc.most_common(25)))
print(f"{'inst%':>6} {'samp%':>6}  file:line  source   [top stalls]")
for (f, ln, src), (inst, samp, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*inst/max(tot_i,1):6.2f} {100*samp/max(tot_s,1):6.2f}  {f}:{ln}  {src}   [{', '.join(f'{k[6:]}={v}' for k, v in st.most_common(3))}]")
