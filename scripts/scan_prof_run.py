import sys, os
sys.path.insert(0,'.')
import torch, numpy as np
import sessionsimilaritysearch_b200 as sss
from sessionsimilaritysearch_b200 import _lib
g=torch.Generator(device='cuda').manual_seed(0)
ix=sss.IndexFlatIP(128, mode="bf16")
for _ in range(20):
    ix.add(torch.randn((250000,128),generator=g,device='cuda'), norm=sss.NORM_UTIL)
q=sss.normalize(torch.randn((1000,128),generator=g,device='cuda'))
for _ in range(3): ix.search(q,100)
lib=_lib.load()
base=[lib.sss_index_stat(ix._h, 16+i) for i in range(8)]
for _ in range(200): ix.search(q,100)
base=[lib.sss_index_stat(ix._h, 16+i) for i in range(8)]
ix.search(q,100)
now=[lib.sss_index_stat(ix._h, 16+i) for i in range(8)]
d=[n-b for n,b in zip(now,base)]
units=d[5]
print('per unit cycles: MMA wait tempty %.0f, wait full %.0f (per unit), issue %.0f ; epi(w0) wait tfull per its unit %.0f; units %d'%(d[0]/units, d[1]/units, d[2]/units, d[3]/(units/ (2*1))*1.0/1, units))

print('effective SM clock during the scan (MMA thread): %.3f GHz'%(d[7]/max(d[6],1)))
