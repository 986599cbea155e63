#!/usr/bin/env python
"""Aggregate an ncu --import-source report by CUDA source line: joins the SASS page of the report (per-instruction
executed counts and stall samples) with nvdisasm's line info for the same kernel.

    python scripts/ncu_by_line.py report.ncu-rep cubin kernel_substr [launch_index] [top]
"""
import csv, io, re, subprocess, sys

rep, cubin, kname = sys.argv[1:4]
launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}; blocks.append(cur); continue
    if cur is None: continue
    if cur["hdr"] is None: cur["hdr"] = r; continue
    if len(r) > 5: cur["rows"].append(r)
blocks = [b for b in blocks if kname in b["name"]]
b = blocks[launch]
h = b["hdr"]
ii, si = h.index("Instructions Executed"), h.index("# Samples")
stall_cols = [(n, h.index(n)) for n in h if n.startswith("stall_") and "Not Issued" not in n]
# line info from nvdisasm
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
lines_by_off, in_fn, cur_line, cur_file = {}, False, None, None
for ln in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln) or re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
    if m:
        in_fn = kname in m.group(1); continue
    if not in_fn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_file, cur_line = m.group(1).split("/")[-1], int(m.group(2)); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m: lines_by_off[int(m.group(1), 16)] = (cur_file, cur_line)
base = int(b["rows"][0][0], 16)
agg, tot_i, tot_s = {}, 0.0, 0.0
for r in b["rows"]:
    off = int(r[0], 16) - base
    key = lines_by_off.get(off, ("?", 0))
    a = agg.setdefault(key, [0.0, 0.0, {}])
    a[0] += float(r[ii] or 0); a[1] += float(r[si] or 0)
    for n, c in stall_cols:
        v = float(r[c] or 0)
        if v: a[2][n] = a[2].get(n, 0) + v
    tot_i += float(r[ii] or 0); tot_s += float(r[si] or 0)
print(f"{b['name']}  launch {launch}: {tot_i:.0f} warp instructions, {tot_s:.0f} samples")
src = {}
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    if f not in src:
        try: src[f] = open(f"/root/repo/evenvizion_b200/csrc/{f}").read().splitlines()
        except OSError: src[f] = []
    text = src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""
    st = ", ".join(f"{n[6:]} {v / max(a[1], 1) * 100:.0f}%" for n, v in sorted(a[2].items(), key=lambda kv: -kv[1])[:3])
    print(f"{a[1] / max(tot_s, 1) * 100:5.1f}% smp {a[0] / max(tot_i, 1) * 100:5.1f}% inst  {f}:{l:<4} {text}   [{st}]")
if len(sys.argv) > 6:
    # extra: aggregate by line ranges "name:lo-hi,name:lo-hi"
    print("--- by region")
    for spec in sys.argv[6].split(","):
        name, rng = spec.split(":"); lo, hi = map(int, rng.split("-"))
        i = sum(a[0] for (f, l), a in agg.items() if lo <= l <= hi)
        s = sum(a[1] for (f, l), a in agg.items() if lo <= l <= hi)
        print(f"{name:24s} {s / max(tot_s, 1) * 100:5.1f}% smp {i / max(tot_i, 1) * 100:5.1f}% inst")
