import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value'); gs=h.index('Grid Size'); bs=h.index('Block Size')
agg=collections.defaultdict(list); grid={}
for r in rows[hi+1:]:
    if len(r)>mv:
        k=r[kn][:70]; agg[k].append(float(r[mv].replace(',',''))); grid[k]=(r[gs],r[bs])
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    print('%-72s n=%4d avg=%9.2f us share=%5.1f%% grid=%s block=%s'%(k,len(v),sum(v)/len(v)/1e3,100*sum(v)/tot,grid[k][0],grid[k][1]))
