#!/usr/bin/env python3
"""Attribute an ncu report's per-SASS-instruction counters of one kernel to the source functions / lines of
trace_impl.cuh.  ncu's own source page needs the original paths; this joins the report's SASS rows with the line table
nvdisasm prints for the same (unchanged) build of libsrt.so.
usage: [SRT_NCU_LAUNCH=k] ncu_by_function.py report.ncu-rep mangled_kernel_name [n_lines]   (k-th profiled launch of the report, default the last)"""
import collections, csv, io, os, pathlib, re, subprocess, sys, tempfile

ROOT = pathlib.Path(__file__).resolve().parents[1]
rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = pathlib.Path(tempfile.mkdtemp())
subprocess.run(["cuobjdump", "-xelf", "all", str(ROOT / "cuda-spectral-ray-tracer_b200" / "libsrt.so")], cwd=tmp, check=True, capture_output=True)
cubin = [p for p in tmp.iterdir() if p.name.startswith(os.environ.get("SRT_NCU_CUBIN", "trace_fast"))][0]
dis = subprocess.run(["nvdisasm", "-g", str(cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith(".text." + kernel + ":")][0]
cur, ins = None, []
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((m.group(2), cur))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# one section per profiled launch: a "Kernel Name" row, a header row, then one row per SASS instruction; SRT_NCU_LAUNCH picks one (default: the last)
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
import os
which = int(os.environ.get("SRT_NCU_LAUNCH", len(starts) - 2))
hdr, data = rows[starts[which] + 1], rows[starts[which] + 2:starts[which + 1]]
H = {h: i for i, h in enumerate(hdr)}
if len(data) != len(ins):
    sys.exit("SASS of the report (%d instructions) and of libsrt.so (%d) differ: rebuild the profiled commit" % (len(data), len(ins)))
src = (ROOT / "cuda-spectral-ray-tracer_b200/csrc/cuda/trace_impl.cuh").read_text().split("\n")
fn, curf = [None] * (len(src) + 2), "?"
for i, l in enumerate(src, 1):
    m = re.match(r"^(?:template.*)?\s*(?:static\s+)?__(?:device|global)__.*?\b(\w+)\s*\(", l)
    if m and not l.startswith(" "):
        curf = m.group(1)
    fn[i] = curf
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
lagg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for (op, c), r in zip(ins, data):
    ie, te, sm = int(r[H["Instructions Executed"]]), int(r[H["Thread Instructions Executed"]]), int(r[H["# Samples"]])
    f = fn[c[1]] if c and c[0] == "trace_impl.cuh" else (c[0] if c else "?")
    if f == "__launch_bounds__":
        f = "(kernel body)"
    for a, k in ((agg, f), (lagg, c or ("?", 0))):
        a[k][0] += ie; a[k][1] += te; a[k][2] += sm
    agg[f][3] += 1
    tot[0] += ie; tot[1] += te; tot[2] += sm
print("total: %.2f G warp instructions, %.1f G thread instructions, %d stall samples" % (tot[0] / 1e9, tot[1] / 1e9, tot[2]))
print("%-28s %8s %8s %8s %6s %5s" % ("function", "winst%", "tinst%", "samp%", "lanes", "sass"))
for f, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-28s %8.1f %8.1f %8.1f %6.1f %5d" % (f, 100 * v[0] / tot[0], 100 * v[1] / tot[1], 100 * v[2] / tot[2], v[1] / max(v[0], 1), v[3]))
print()
for k, v in sorted(lagg.items(), key=lambda kv: -kv[1][2])[:top]:
    s = src[k[1] - 1].strip()[:100] if k[0] == "trace_impl.cuh" else ""
    print("%-20s %5d samp%%%5.1f winst%%%5.1f lanes %4.1f | %s" % (k[0][:20], k[1], 100 * v[2] / tot[2], 100 * v[0] / tot[0], v[1] / max(v[0], 1), s))
