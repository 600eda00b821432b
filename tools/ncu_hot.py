"""Summarise an ncu report's source page per CUDA source line: stall samples + dominant stall reasons.
usage: python tools/ncu_hot.py report.ncu-rep [topN] [kernel-regex]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
kf = ["-k", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + kf, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = "?"; hdr = None; cur = None
agg = collections.defaultdict(lambda: collections.Counter()); src = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Name":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Kernel Name":
        continue
    if r[0] == "Line No":
        hdr = r; ci = {}
        for k, h in enumerate(hdr):
            ci.setdefault(h, k)
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0].strip():
        cur = (fname, int(r[0])); src[cur] = r[1].strip()
    s = r[ci["# Samples"]]
    if cur and s.isdigit() and int(s):
        agg[cur]["n"] += int(s)
        for st in stalls:
            v = r[ci[st]]
            if v.isdigit() and int(v):
                agg[cur][st[6:]] += int(v)
tot = sum(a["n"] for a in agg.values())
mix = collections.Counter()
for a in agg.values():
    for k, v in a.items():
        if k != "n":
            mix[k] += v
print("total samples", tot, "| stall mix:", ", ".join("%s %.0f%%" % (k, 100 * v / max(tot, 1)) for k, v in mix.most_common(7)))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["n"])[:topn]:
    top = [(k, v) for k, v in a.most_common(4) if k != "n"][:3]
    print("%5.1f%% %s:%d  %-90s %s" % (100 * a["n"] / tot, key[0], key[1], src.get(key, "")[:90], " ".join("%s:%d" % kv for kv in top)))
