"""Stall samples of one kernel per CUDA source line: joins an `ncu --page source --csv` dump (per-SASS-instruction
samples) with `nvdisasm -g` line info of the same build.
    python profiles/tools/line_hot.py <lib.so> <kernel substring> <ncu source csv> [top N]
Inlined code is attributed to the innermost source line (the line of the inlined callee)."""
import csv, os, re, subprocess, sys, tempfile, collections

lib, kern, src_csv = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# walk the function: map instruction offset -> (file, line)
inside, cur, off2line = False, ("?", 0), {}
for l in dis:
    if l.startswith("//-") and ".text." in l:
        inside = kern in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        off2line[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {n: i for i, n in enumerate(hdr)}
seen, recs = set(), []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] in seen or not r[0].strip():
        continue
    seen.add(r[0])
    try:
        recs.append((int(r[0], 16), float(r[ci["# Samples"]] or 0), float(r[ci["Instructions Executed"]] or 0)))
    except ValueError:
        pass
base = min(a for a, _, _ in recs)
per = collections.defaultdict(lambda: [0.0, 0.0, 0])
tot = sum(s for _, s, _ in recs)
for a, s, ex in recs:
    ln = off2line.get(a - base, (("?", 0), ""))[0]
    per[ln][0] += s; per[ln][1] += ex; per[ln][2] += 1
srcs = {}
def text(f, n):
    if f not in srcs:
        for root in ("rl_rocket_6dof_b200/csrc", "include"):
            p = os.path.join(root, f)
            if os.path.exists(p):
                srcs[f] = open(p).read().splitlines(); break
        else:
            srcs[f] = []
    L = srcs[f]
    return L[n - 1].strip()[:100] if 0 < n <= len(L) else ""
print(f"{kern}: {tot:.0f} samples, {len(recs)} SASS instructions")
for (f, n), (s, ex, cnt) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * s / tot:5.1f}%  {f}:{n:<5d} ({cnt:3d} instr)  {text(f, n)}")
