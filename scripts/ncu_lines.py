"""Aggregate an ncu report's executed instructions / stall samples by source line.
usage: python scripts/ncu_lines.py report.ncu-rep kernel_substring [top]"""
import collections, csv, glob, io, os, re, subprocess, sys, tempfile
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "multirobot_pathplanning_benchmark_b200", "libmrb200.so")], cwd=tmp, capture_output=True)
dis = None
for c in glob.glob(os.path.join(tmp, "*.cubin")):
    o = subprocess.run(["nvdisasm", "-g", "-c", c], capture_output=True, text=True).stdout
    if kern in o:
        dis = o
        break
addr2line, cur, infunc = {}, ("?", 0), False
for l in dis.split("\n"):
    if ".section" in l or l.startswith(".text."):
        infunc = kern in l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m and infunc:
        addr2line[int(m.group(1), 16)] = cur
base = int(data[0][0], 16)
agg, samp, tot = collections.Counter(), collections.Counter(), 0
stl = collections.defaultdict(collections.Counter)
for r in data:
    try:
        a = int(r[0], 16) - base
    except ValueError:
        continue
    fl = addr2line.get(a, ("?", 0))
    n = int(r[ia] or 0)
    agg[fl] += n
    samp[fl] += int(r[isamp] or 0)
    for i in stall_cols:
        stl[fl][hdr[i][6:]] += int(r[i] or 0)
    tot += n
print("total warp instructions", tot, "sass lines", len(data))
srcs = {}
for fl, n in agg.most_common(top):
    path = glob.glob(os.path.join(root, "**", fl[0]), recursive=True)
    txt = ""
    if path:
        srcs.setdefault(path[0], open(path[0]).read().split("\n"))
        if 0 < fl[1] <= len(srcs[path[0]]):
            txt = srcs[path[0]][fl[1] - 1].strip()[:90]
    why = " ".join(f"{k}:{v * 100 // max(samp[fl], 1)}" for k, v in stl[fl].most_common(3))
    print(f"{fl[0]:20s} {fl[1]:4d} {n / tot * 100:5.1f}%  smp {samp[fl]:6d}  [{why}]  {txt[:60]}")
