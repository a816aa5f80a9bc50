"""Split an ncu source-page export (cuda,sass) into the regions between BAR.SYNC instructions (the kernel's stages) in
SASS address order: warp-instructions executed, share, stall samples, top opcodes per region."""
import csv, sys, collections
rows = []
hdr = None
for r in csv.reader(open(sys.argv[1])):
    if not r: continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] != "" or len(r) != len(hdr) or not r[2].startswith("0x"): continue
    d = dict(list(zip(hdr, r))[4:])
    try:
        rows.append((int(r[2], 16), r[3].strip(), int(d.get("Instructions Executed") or 0), int(d.get("# Samples") or 0), d))
    except ValueError:
        pass
seen = {}
for a in rows: seen[a[0]] = a
rows = [seen[k] for k in sorted(seen)]
tot = sum(r[2] for r in rows); tots = sum(r[3] for r in rows)
reg = 0; acc = collections.defaultdict(lambda: [0, 0, collections.Counter(), collections.Counter(), 0])
for addr, sass, inst, samp, d in rows:
    a = acc[reg]
    a[0] += inst; a[1] += samp; a[4] += 1
    t = sass.split(); op = (t[1] if t[0].startswith("@") else t[0]).rstrip(";").split(".")[0]
    a[2][op] += inst
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0", "-"): a[3][k[6:]] += int(v)
    if "BAR.SYNC" in sass: reg += 1
px = float(sys.argv[2]) if len(sys.argv) > 2 else 0
print(f"total warp-inst {tot}  samples {tots}" + (f"  lane-inst/px {tot*32/px:.1f}" if px else ""))
for k in sorted(acc):
    a = acc[k]
    print(f"region {k}: sass {a[4]:5d}  inst {100*a[0]/tot:5.1f}%  samp {100*a[1]/max(tots,1):5.1f}%" + (f"  lane-inst/px {a[0]*32/px:5.2f}" if px else "")
          + "  ops: " + ", ".join(f"{o}={100*v/a[0]:.0f}%" for o, v in a[2].most_common(8)) + "  stalls: " + ", ".join(f"{o}={v}" for o, v in a[3].most_common(4)))
