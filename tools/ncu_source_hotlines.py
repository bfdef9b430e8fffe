import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
top=int(sys.argv[2]) if len(sys.argv)>2 else 40
sections=[]; cur=None
for r in rows:
    if r and r[0]=='File Path':
        cur={'file':r[1],'rows':[]}; sections.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
seen=set()
grand=0
res=[]
for s in sections:
    if s['file'] in seen: continue
    seen.add(s['file'])
    hdr=None
    tot=collections.Counter(); src={}; stall=collections.Counter()
    for r in s['rows']:
        if r and r[0]=='Line No': hdr=r; continue
        if hdr is None or len(r)<len(hdr): continue
        try: ln=int(r[0])
        except: continue
        ie=hdr.index('Instructions Executed'); ss=hdr.index('# Samples')
        try: v=float(r[ie])
        except: continue
        tot[ln]+=v; src[ln]=r[1]
        try: stall[ln]+=float(r[ss])
        except: pass
    res.append((s['file'],tot,src,stall)); grand+=sum(tot.values())
print('grand total inst',grand)
allst=sum(sum(st.values()) for _,_,_,st in res)
for f,tot,src,stall in res:
    T=sum(tot.values())
    if T<grand*0.01: continue
    print('==',f,'inst %.0f (%.1f%%)'%(T,100*T/grand))
    for ln,v in tot.most_common(top):
        print('%5d inst %5.2f%% samp %5.2f%% %s'%(ln,100*v/grand,100*stall[ln]/max(allst,1),src[ln][:100]))
