import csv,sys,collections
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ii], r[ki].replace('adrt_b200::','').replace('(anonymous namespace)::','').replace('<unnamed>::','')[:70]),{})[r[mi]]=float(r[vi].replace(',',''))
items=list(d.items())
for (i,k),m in items[len(items)//2:]:
    print('%8.1f us  rd %7.1f MB  wr %7.1f MB  %s' % (m.get('gpu__time_duration.sum',0)/1000.0, m.get('dram__bytes_read.sum',0)/1e6, m.get('dram__bytes_write.sum',0)/1e6, k))
