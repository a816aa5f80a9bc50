"""Static view of one kernel's SASS (no GPU needed): every loop (backward branch) with its instruction count, the
alu / fma / lsu split and the source lines it comes from.  Used to budget lane-instructions per pixel before spending
GPU time: instr/px = sum over loops (instructions x trips) x 32 / pixels.

usage: sass_static.py [lib.so] [kernel-name-substring]   (default: cudacam_b200/libb200canny.so k_stencil_marchILi3)
"""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "cudacam_b200", "libb200canny.so")
want = sys.argv[2] if len(sys.argv) > 2 else "k_stencil_marchILi3"
FMA = ("FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "HADD2", "HMUL2", "FHFMA", "IDP", "FHADD", "HSETP2", "HSET2")
LSU = ("LDG", "STG", "LDS", "STS", "ATOMS", "ATOMG", "RED", "SHFL", "LDC", "LDCU", "LD", "ST", "BAR", "VOTE", "CCTL")
def pipe(op):
    b = op.split(".")[0]
    if b in FMA: return "fma"
    if b in LSU: return "lsu"
    if b in ("BRA", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "NOP", "CALL", "RET", "BREAK", "NANOSLEEP"): return "ctl"
    if b in ("I2F", "F2I", "MUFU", "I2FP", "F2FP", "POPC", "FLO", "BREV"): return "xu"
    return "alu"
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, check=True, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin") and "synth" not in f][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
lines = txt.split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and want in l)
ins = []   # (idx, label-or-None, op, text, srcline)
labels = {}
cur = None
for l in lines[start + 1:]:
    if l.startswith("\t.section") or l.startswith(".text."): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"^(\.L_x_\d+):", l)
    if m: labels[m.group(1)] = len(ins); continue
    m = re.match(r"^\s+/\*[0-9a-f]+\*/\s+(.*?);", l)
    if m:
        t = m.group(1).split()
        op = t[1] if t[0].startswith("@") else t[0]
        ins.append((len(ins), op, m.group(1), cur))
loops = []
for i, op, text, src in ins:
    if op.startswith("BRA"):
        m = re.search(r"(\.L_x_\d+)", text)
        if m and m.group(1) in labels and labels[m.group(1)] <= i: loops.append((labels[m.group(1)], i))
loops.sort(key=lambda ab: (ab[0], -ab[1]))
print(f"{want}: {len(ins)} instructions = {len(ins) * 16 / 1024:.1f} KB, {len(loops)} loops")
def own(a, b):
    inner = [(x, y) for x, y in loops if a <= x and y <= b and (x, y) != (a, b)]
    return [k for k in range(a, b + 1) if not any(x <= k <= y for x, y in inner)]
for a, b in loops:
    ks = own(a, b)
    pc = collections.Counter(pipe(ins[k][1]) for k in ks)
    srcs = collections.Counter(ins[k][3] for k in ks if ins[k][3])
    lo = min((s[1] for s in srcs if s[0].startswith("k_stencil_march")), default=0); hi = max((s[1] for s in srcs if s[0].startswith("k_stencil_march")), default=0)
    ops = collections.Counter(ins[k][1].split(".")[0] for k in ks)
    print(f"loop [{a:5d},{b:5d}] total {b - a + 1:5d} own {len(ks):5d}  alu {pc['alu']:4d} fma {pc['fma']:4d} lsu {pc['lsu']:4d} xu {pc['xu']:3d} ctl {pc['ctl']:3d}  lines {lo}-{hi}  top: " + ", ".join(f"{o}={n}" for o, n in ops.most_common(8)))
if "--dump" in sys.argv:
    a, b = map(int, sys.argv[sys.argv.index("--dump") + 1].split(","))
    for k in range(a, b + 1): print(f"{k:5d} {pipe(ins[k][1]):3s} {str(ins[k][3]):32s} {ins[k][2]}")
