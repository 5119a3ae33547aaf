#!/usr/bin/env python
"""Executed warp-instructions and stall samples of one profiled kernel by CUDA source line.

  tools/ncu_by_line.py REPORT.ncu-rep MANGLED_KERNEL_NAME [units]

Joins `ncu --page source --csv` (SASS level: one row per instruction with its address) with the line table of the
built library (`nvdisasm -g` of the cubin that holds the kernel).  `units` scales the counts (e.g. items per launch).
The library must be the build that was profiled."""
import collections, csv, glob, io, os, re, subprocess, sys, tempfile

rep, kern = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "bias.jl_b200", "csrc", "libbias_cuda.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
line_of = {}
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    if kern not in txt:
        continue
    inside, cur = False, None
    for ln in txt.split("\n"):
        if ".section" in ln and ".text." in ln:
            inside = kern in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hdr = rows[1]
ia, iad, ist = hdr.index("Instructions Executed"), hdr.index("Address"), hdr.index("# Samples")
base = int(rows[2][iad], 16)
inst, samp = collections.Counter(), collections.Counter()
for r in rows[2:]:
    try:
        n, off, s = int(r[ia]), int(r[iad], 16) - base, int(r[ist])
    except (ValueError, IndexError):
        continue
    inst[line_of.get(off)] += n
    samp[line_of.get(off)] += s
src = {}
tot, ts = sum(inst.values()), max(1, sum(samp.values()))
print(f"total {tot / units:.1f} warp-instructions per unit, {ts} stall samples")
for key, n in inst.most_common(40):
    f, l = key if key else ("?", 0)
    if f not in src:
        path = os.path.join(root, "bias.jl_b200", "csrc", f)
        src[f] = open(path).read().split("\n") if os.path.exists(path) else []
    text = src[f][l - 1].strip()[:100] if 0 < l <= len(src[f]) else ""
    print(f"{n / units:9.1f} {100 * samp[key] / ts:5.1f}%  {f}:{l}  {text}")
