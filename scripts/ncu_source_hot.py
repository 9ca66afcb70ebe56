"""Per-CUDA-line instruction and stall-sample totals from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`:
    ncu -i rep --page source --print-source cuda,sass --csv --kernel-id :::N | python scripts/ncu_source_hot.py [top]"""
import csv
import sys

top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
cur_file, hdr, out = "", None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0]:
        continue
    try:
        ie = int(r[hdr.index("Instructions Executed")])
        s = int(r[hdr.index("# Samples")])
    except ValueError:
        continue
    out.append((ie, s, cur_file, r[0], r[1][:105]))
tot = sum(o[0] for o in out) or 1
tots = sum(o[1] for o in out) or 1
print("total warp instructions %d, samples %d" % (tot, tots))
print("-- by instructions")
for o in sorted(out, reverse=True)[:top]:
    print("%9d %5.1f%% smp %5.1f%% | %s:%s | %s" % (o[0], 100 * o[0] / tot, 100 * o[1] / tots, o[2], o[3], o[4]))
print("-- by stall samples")
for o in sorted(out, key=lambda x: -x[1])[:top]:
    print("%9d %5.1f%% smp %5.1f%% | %s:%s | %s" % (o[0], 100 * o[0] / tot, 100 * o[1] / tots, o[2], o[3], o[4]))
