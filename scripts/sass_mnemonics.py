"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMA / cp.async use:
cuobjdump -sass sessionsimilaritysearch_b200/libsss_b200.so | python scripts/sass_mnemonics.py"""
import collections
import re
import sys

cur, cnt = None, collections.defaultdict(collections.Counter)
for line in sys.stdin:
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and cur:
        cnt[cur][m.group(1).split('.')[0]] += 1
keys = ('UTCHMMA', 'UTCQMMA', 'UTMALDG', 'LDTM', 'STTM', 'LDGSTS', 'POPC', 'HMMA', 'HGMMA')
for f, c in sorted(cnt.items()):
    if any(c[k] for k in keys[:6]) or 'hamming' in f:
        short = re.sub(r'_ZN3sss\d+_GLOBAL__N__[0-9a-f_]+cu_[0-9a-f]+', '', f)[:80]
        print(short, {k: c[k] for k in keys if c[k]})
print('HMMA + HGMMA anywhere:', sum(c['HMMA'] + c['HGMMA'] for c in cnt.values()))
