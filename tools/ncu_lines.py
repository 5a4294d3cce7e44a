#!/usr/bin/env python
"""Attribute ncu warp-stall samples to CUDA source lines.

usage: tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [LIB.so] [--top N]
Joins `ncu --page source --csv` (per-SASS-instruction samples) with `nvdisasm -g` line info of the library's cubin.
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, kre = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else "axial_vs_b200/libaxvs.so"
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 30

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kname = rows[0][1]
hdr = rows[1]
ia, iss, isrc = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
inst = []
seen = set()
for r in rows[2:]:
    if len(r) <= iss or r[ia] in seen:
        continue
    seen.add(r[ia])
    try:
        inst.append((int(r[ia], 16), int(r[iss] or 0), r[isrc], {hdr[i][6:]: int(r[i]) for i in stall_cols if r[i] not in ("", "0")}))
    except ValueError:
        pass
base = min(a for a, *_ in inst)

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
base_name = re.sub(r"[<(].*", "", kname.replace("void ", "")).split("::")[-1]
targ = re.search(r"<\(int\)(\d+)>", kname)
mangled = base_name + (f"ILi{targ.group(1)}E" if targ else "")
line_of = {}
cur = None
infunc = False
for ln in dis.splitlines():
    if ln.startswith(".text."):
        infunc = mangled in ln
        continue
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.search(r"/\*([0-9a-f]{4,})\*/", ln)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur

agg = collections.defaultdict(lambda: [0, collections.Counter()])
tot = 0
for a, s, src, st in inst:
    key = line_of.get(a - base, ("?", 0))
    agg[key][0] += s
    agg[key][1].update(st)
    tot += s
print(f"{kname}: {tot} samples")
srcs = {}
for (f, l), (s, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        p = os.path.join("axial_vs_b200/csrc", f)
        srcs[f] = open(p).read().splitlines() if os.path.isfile(p) else []
    text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
    print(f"{s:6d} {100.0 * s / max(tot, 1):5.1f}%  {f}:{l:<4d} {dict(st.most_common(2))}  | {text}")
