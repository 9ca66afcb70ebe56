"""Summarise `ncu -i X.ncu-rep --page raw --csv` (stdin): one block per captured launch with the metrics the roofline
discussion in profiles/ uses."""
import csv
import sys

rows = [r for r in csv.reader(l for l in sys.stdin if l.startswith('"'))]
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active']
idx = [(w, hdr.index(w)) for w in want if w in hdr]
tens = [(h, i) for i, h in enumerate(hdr) if 'tensor' in h and '.avg.pct' in h]
units = rows[1]
for r in rows[2:]:
    print("--")
    for w, i in idx:
        print("  %-62s %s %s" % (w, r[i][:70], units[i]))
    for h, i in tens:
        if r[i] in ('0', '0.000000', ''):
            continue
        print("  %-62s %s %s" % (h, r[i], units[i]))
