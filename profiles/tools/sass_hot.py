"""Per-SASS-instruction view of an `ncu --page source --csv` dump: executions per warp, stall samples, opcode mix by
region.  Usage: python profiles/tools/sass_hot.py <source.csv> <warps_launched> [--list]"""
import csv, sys, collections, re

path, warps = sys.argv[1], float(sys.argv[2])
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
ins = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr) or not r[0].strip():
        continue
    try:
        ex = float(r[col["Instructions Executed"]])
    except ValueError:
        continue
    src = r[col["Source"]].strip()
    m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)", src)
    op = m.group(2) if m else "?"
    ins.append(dict(addr=r[0], src=src, op=op, ex=ex, samples=float(r[col["# Samples"]] or 0),
                    wait=float(r[col["stall_wait"]] or 0), ssb=float(r[col["stall_short_sb"]] or 0),
                    lsb=float(r[col["stall_long_sb"]] or 0), math=float(r[col["stall_math"]] or 0),
                    thr=float(r[col["Avg. Threads Executed"]] or 0)))
seen=set(); ins=[i for i in ins if not (i["addr"] in seen or seen.add(i["addr"]))]
tot = sum(i["ex"] for i in ins)
tots = sum(i["samples"] for i in ins)
print(f"static {len(ins)} instr; executed {tot:.0f} = {tot / warps:.1f} per warp; samples {tots:.0f}")
by = collections.Counter()
for i in ins:
    by[i["op"]] += i["ex"]
print("opcode: executions per warp")
for op, c in by.most_common(30):
    print(f"  {op:8s} {c / warps:8.1f}  {100 * c / tot:5.1f}%")
fp64 = sum(c for op, c in by.items() if op in ("DFMA", "DMUL", "DADD", "DSETP"))
print(f"FP64 instr per warp {fp64 / warps:.1f} ({100 * fp64 / tot:.1f}%)")
if "--list" in sys.argv:
    for i in ins:
        if i["ex"] > 0:
            print(f'{i["addr"][-5:]} {i["ex"] / warps:7.2f} s={i["samples"]:5.0f} w={i["wait"]:4.0f} ss={i["ssb"]:4.0f} ls={i["lsb"]:4.0f} m={i["math"]:4.0f} t={i["thr"]:4.1f} {i["src"][:90]}')
