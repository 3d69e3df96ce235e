import sys,re,collections
for tag in ("cl1","cl0"):
    g=collections.OrderedDict()
    for l in open(sys.argv[1]):
        if not l.startswith(tag+" "): continue
        m=re.search(r'kind=(\d) cin=\s*(\d+) cout=\s*(\d+) k=\s*(\d+) d=(\d) L=\s*(\d+)\s+([\d.]+) ms',l)
        if not m: continue
        kind,cin,cout,k,d,L,ms=m.groups(); g.setdefault((kind,cin,cout,k),[]).append(float(ms))
    tot=0
    print("==", tag)
    for k,v in g.items():
        print(k, 'n=%d'%len(v), 'sum=%.3f'%sum(v), ' '.join('%.3f'%x for x in v)); tot+=sum(v)
    print('total',tot)
