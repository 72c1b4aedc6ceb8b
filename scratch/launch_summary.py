import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
for i,r in enumerate(rows):
    if r and r[0]=='ID': hdr=r; start=i+1; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.OrderedDict()
for r in rows[start:]:
    if len(r)<=vi: continue
    n=r[ki][:70]; v=float(r[vi].replace(',',''))
    if r[ui]=='ns': v/=1000
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
print('| kernel | launches | total us | share |\n|---|---:|---:|---:|')
for n,a in sorted(agg.items(), key=lambda x:-x[1][1]): print(f'| `{n}` | {a[0]} | {a[1]:.1f} | {a[1]/tot:.3f} |')
print(f'\nTotal {tot:.1f} us')
