import csv, sys
def sections(path):
    rows=list(csv.reader(open(path)))
    print(rows[0][1][:100])
    secs=[]; cur=None
    for r in rows:
        if r and r[0]=='Address': cur=[r]; secs.append(cur)
        elif cur is not None and len(r)>5: cur.append(r)
    return secs
def f(x):
    try: return float(x)
    except: return 0.0
def top(path, n=40, which=0):
    secs=sections(path)
    print('sections', [len(s) for s in secs])
    s=secs[which]; hdr=s[0]; body=s[1:]
    iS=hdr.index('# Samples'); iSrc=hdr.index('Source'); iEx=hdr.index('Instructions Executed')
    stall=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot=sum(f(r[iS]) for r in body); totex=sum(f(r[iEx]) for r in body)
    print('total samples', tot, 'instr executed', totex, 'lines', len(body))
    agg={}
    for r in body:
        for i in stall: agg[hdr[i]]=agg.get(hdr[i],0)+f(r[i])
    print(sorted(((int(v),k) for k,v in agg.items()), reverse=True)[:8])
    idx=sorted(range(len(body)), key=lambda k:-f(body[k][iS]))[:n]
    for k in sorted(idx):
        r=body[k]
        st=sorted(((f(r[i]),hdr[i]) for i in stall), reverse=True)[:2]
        print('%5d %5.1f%% ex=%9s  %-78s %s'%(k,100*f(r[iS])/tot, r[iEx], r[iSrc][:78], ' '.join('%s=%d'%(h[6:],v) for v,h in st if v>0)))
top(sys.argv[1], int(sys.argv[2]) if len(sys.argv)>2 else 40, int(sys.argv[3]) if len(sys.argv)>3 else 0)
