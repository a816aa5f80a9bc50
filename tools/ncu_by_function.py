"""Attributes the executed warp-instructions of one ncu capture (--set full --import-source on) to the FUNCTIONS of the
kernel's source file (by the line ranges of the function definitions; inlined helpers from other headers are listed by
file) and prints lane-instructions per pixel, the pipe split and the top stall samples per function, then the heaviest
source lines.   usage: ncu_by_function.py report.ncu-rep [pixels_per_launch] [source.cuh]"""
import collections, csv, io, os, re, subprocess, sys
rep = sys.argv[1]
px = float(sys.argv[2]) if len(sys.argv) > 2 else 64 * 1920 * 1080
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
srcname = sys.argv[3] if len(sys.argv) > 3 else "k_stencil_march.cuh"
src = open(os.path.join(ROOT, "cudacam_b200", "csrc", srcname)).read().split("\n")
marks = []
for i, l in enumerate(src, 1):
    m = re.match(r"(?:template <[^>]*>\s*)?(?:__device__ __forceinline__|__global__|inline|static inline)\s+[\w:<> ]*?\b(\w+)\s*\(", l)
    if m and not l.startswith(" "): marks.append((i, m.group(1)))
    m = re.search(r"// ---- (.*?) -{3,}", l)
    if m and l.startswith("  "): marks.append((i, "  " + m.group(1)[:40]))
def fn(ln):
    s = "?"
    for l, n in marks:
        if l <= ln: s = n
    return s
FMA = ("FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "HADD2", "HMUL2", "FHFMA", "IDP", "HSETP2", "HSET2")
LSU = ("LDG", "STG", "LDS", "STS", "ATOMS", "ATOMG", "RED", "SHFL", "LDC", "LDCU", "LD", "ST", "BAR", "VOTE", "SYNCS")
def pipe(op):
    b = op.split(".")[0]
    if b in FMA: return "fma"
    if b in LSU: return "lsu"
    if b.startswith("U") and b not in ("UNPACK",): return "uni"
    if b in ("BRA", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "NOP", "CALL", "RET", "BREAK", "NANOSLEEP", "BRX"): return "ctl"
    if b in ("I2F", "F2I", "MUFU", "I2FP", "F2FP", "POPC", "FLO", "BREV"): return "xu"
    return "alu"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
hdr = None; rows = []; curfile = "?"; curline = ("?", 0)
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
    stalls = {k[len("stall_"):]: int(v) for k, v in d.items() if k.startswith("stall_") and v and v.isdigit() and int(v)}
    rows.append((int(r[2], 16), curline, r[3].strip(), inst, samp, stalls))
seen = {a[0]: a for a in rows}
rows = [seen[k] for k in sorted(seen)]
agg = collections.Counter(); sm = collections.Counter(); pp = collections.defaultdict(collections.Counter); st = collections.defaultdict(collections.Counter)
lines = collections.Counter(); lsm = collections.Counter(); tot = 0
for addr, (f, ln), sass, inst, samp, stalls in rows:
    key = fn(ln) if f == srcname else f
    t = sass.split(); op = (t[1] if t[0].startswith("@") else t[0]).rstrip(";")
    agg[key] += inst; sm[key] += samp; tot += inst; pp[key][pipe(op)] += inst
    for k, v in stalls.items(): st[key][k] += v
    lines[(f, ln)] += inst; lsm[(f, ln)] += samp
ts = sum(sm.values())
print(f"total {tot * 32 / px:.2f} lane-instr/px, {tot} warp-instructions, {ts} samples")
allp = collections.Counter()
for k in pp: allp.update(pp[k])
print("pipes: " + ", ".join(f"{k}={100 * v / tot:.1f}%" for k, v in allp.most_common()))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    if v == 0: continue
    print(f"{k:28s} {100 * v / tot:5.1f}%  {v * 32 / px:6.2f}/px  samp {100 * sm[k] / max(ts, 1):5.1f}%  " + " ".join(f"{a}={100 * b / v:.0f}" for a, b in pp[k].most_common()) + "  | " +
          ", ".join(f"{a}={b}" for a, b in st[k].most_common(4)))
print("\nheaviest lines (inst%  samp%)")
for (f, ln), v in lines.most_common(28):
    text = src[ln - 1].strip()[:110] if f == srcname and 0 < ln <= len(src) else ""
    print(f"{100 * v / tot:5.2f} {100 * lsm[(f, ln)] / max(ts, 1):5.2f}  {f}:{ln}  {text}")
