import csv, sys
f = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/r2h_enc_launches.csv'
lines = [l for l in open(f) if l.startswith('"')]
r = list(csv.reader(lines))
hdr = r[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); gi = hdr.index('Grid Size')
rows = [(x[ki].split('(')[0].replace('void ', '').replace('sss::', '').replace('<unnamed>::', '')[:30], float(x[vi].replace(',', '')) / 1000.0, x[gi]) for x in r[1:]]
idx = [i for i, (n, _, _) in enumerate(rows) if n.startswith('ingest')]
seq = rows[idx[-1]:]
seq = seq[:next((i for i, (n, _, _) in enumerate(seq[1:], 1) if n.startswith('ingest')), len(seq))]
print('launches', len(seq), 'sum us', round(sum(t for _, t, _ in seq), 1))
for n, t, g in seq:
    print('%-32s %8.1f  grid %s' % (n, t, g))
