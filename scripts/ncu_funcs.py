"""Aggregate an ncu report's executed instructions / stall samples by SASS sub-function of one kernel.
usage: python scripts/ncu_funcs.py report.ncu-rep kernel_substring"""
import collections, csv, glob, io, os, re, subprocess, sys, tempfile
rep, kern = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "multirobot_pathplanning_benchmark_b200", "libmrb200.so")], cwd=tmp, capture_output=True)
dis = None
for c in glob.glob(os.path.join(tmp, "*.cubin")):
    o = subprocess.run(["nvdisasm", "-c", c], capture_output=True, text=True).stdout
    if kern in o:
        dis = o
        break
addr2fn, infunc, cur = {}, False, "main"
for l in dis.split("\n"):
    if l.startswith(".text."):
        infunc = kern in l
        cur = "main"
    m = re.match(r"^(\$\S+):", l)
    if m and infunc:
        cur = subprocess.run(["c++filt", m.group(1).split("$")[-1]], capture_output=True, text=True).stdout.strip()[:70]
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m and infunc:
        addr2fn[int(m.group(1), 16)] = cur
base = int(data[0][0], 16)
agg, samp, st, cnt = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter()
for r in data:
    try:
        a = int(r[0], 16) - base
    except ValueError:
        continue
    fn = addr2fn.get(a, "?")
    agg[fn] += int(r[ia] or 0)
    samp[fn] += int(r[isamp] or 0)
    cnt[fn] += 1
    for i in stalls:
        st[fn][hdr[i]] += int(r[i] or 0)
tot, ts = sum(agg.values()), sum(samp.values())
print(f"total warp instructions {tot}, samples {ts}")
for fn, n in agg.most_common():
    top = ", ".join(f"{k[6:]} {v / max(samp[fn], 1) * 100:.0f}%" for k, v in st[fn].most_common(4))
    print(f"{fn:70s} sass {cnt[fn]:5d} inst {n / tot * 100:5.1f}%  samples {samp[fn] / ts * 100:5.1f}%  [{top}]")
