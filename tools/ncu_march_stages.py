"""Attribute the executed warp-instructions of one k_stencil_march capture to the kernel's stages (by the source line
markers `// ---- stage X` in k_stencil_march.cuh; inlined helpers inherit the stage of the surrounding code).
usage: ncu_march_stages.py report.ncu-rep [pixels_per_launch] [--dump STAGE]"""
import collections, csv, io, re, subprocess, sys, os
rep = sys.argv[1]
px = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else 64 * 1920 * 1080
dump = sys.argv[sys.argv.index("--dump") + 1] if "--dump" in sys.argv else None
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "cudacam_b200", "csrc", "k_stencil_march.cuh")).read().split("\n")
marks = []   # (line, name)
for i, l in enumerate(src, 1):
    m = re.search(r"// ---- (prologue|stages? [A-Z/ ]+)", l)
    if m: marks.append((i, m.group(1).replace("stages ", "").replace("stage ", "")))
    m = re.match(r"__device__ __forceinline__ void (m_\w+)|__global__ void .*(k_stencil_march)", l)
    if m: marks.append((i, m.group(1) or "setup"))
marks.sort()
def stage(ln):
    s = "?"
    for l, n in marks:
        if l <= ln: s = n
    return s
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
hdr = None; rows = []
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": curfile = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    if r[0].isdigit(): curline = (curfile, int(r[0])); continue
    if r[0] != "" or len(r) != len(hdr) or not r[2].startswith("0x"): continue
    d = dict(list(zip(hdr, r))[4:])
    try: inst = int(d.get("Instructions Executed") or 0); samp = int(d.get("# Samples") or 0)
    except ValueError: continue
    rows.append((int(r[2], 16), curline, r[3].strip(), inst, samp))
seen = {a[0]: a for a in rows}
rows = [seen[k] for k in sorted(seen)]
agg = collections.Counter(); sm = collections.Counter(); opc = collections.defaultdict(collections.Counter); tot = 0; last = "?"
for addr, (f, ln), sass, inst, samp in rows:
    if f == "k_stencil_march.cuh": last = stage(ln)
    st = last
    agg[st] += inst; sm[st] += samp; tot += inst
    t = sass.split(); op = (t[1] if t[0].startswith("@") else t[0]).rstrip(";")
    opc[st][op] += inst
    if dump and st.strip() == dump: print(f"{addr & 0xffffff:6x} {inst*32/px:6.3f} {samp:4d} {f[-12:]}:{ln:<4d} {sass[:100]}")
ts = sum(sm.values())
print(f"total {tot*32/px:.2f} lane-instr/px")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f"{k:14s} {100*v/tot:5.1f}%  {v*32/px:6.2f}/px  samp {100*sm[k]/max(ts,1):5.1f}%  ", ", ".join(f"{o}={100*c/v:.0f}%" for o, c in opc[k].most_common(10)))
