"""Top CUDA source lines by warp-stall samples, from
  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv
  python tools/ncu_source_lines.py src.csv [top]
"""
import csv, sys, collections
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cur, files, blocks = None, None, collections.OrderedDict()
hdr = None
for r in csv.reader(open(sys.argv[1])):
    if not r:
        continue
    if r[0] == "File Path":
        files = r[1]; continue
    if r[0] == "Function Name":
        cur = r[1][:64]; blocks.setdefault(cur, collections.defaultdict(lambda: [0, 0, ""])); continue
    if r[0] == "Line No":
        hdr = r; si = r.index("# Samples"); ii = r.index("Instructions Executed"); continue
    if cur is None or hdr is None or not r[0].isdigit() or len(r) <= ii:
        continue
    key = (files.rsplit("/", 1)[-1], int(r[0]))
    b = blocks[cur][key]
    try:
        b[0] += int(r[si] or 0); b[1] += int(r[ii] or 0); b[2] = r[1]
    except ValueError:
        pass
for name, d in blocks.items():
    tot = sum(v[0] for v in d.values()); toti = sum(v[1] for v in d.values())
    print("=====", name, "samples", tot, "warp-inst", toti)
    for (f, ln), (s, i, src) in sorted(d.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100*s/max(tot,1):5.1f}% smp {100*i/max(toti,1):5.1f}% inst  {f}:{ln}: {src.strip()[:100]}")
