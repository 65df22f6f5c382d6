import csv,collections,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]
ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); idi=hdr.index('ID')
per=collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[idi],{'k':r[ki]})[r[mi]]=float(r[vi].replace(',',''))
ids=list(per); half=len(ids)//2
agg=collections.defaultdict(lambda:[0,0.0,0.0])
for i in ids[half:]:
    d=per[i]; k=d['k'].split('(')[0].split('::')[-1]
    agg[k][0]+=1; agg[k][1]+=d['gpu__time_duration.sum']; agg[k][2]+=d.get('dram__bytes_read.sum',0)+d.get('dram__bytes_write.sum',0)
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(),key=lambda kv:-kv[1][1]): print('%-34s n=%3d  %.3f ms (%.1f%%)  dram %.1f MB'%(k,v[0],v[1]/1e6,100*v[1]/tot,v[2]/1e6))
print('total ms',tot/1e6)
