import csv,sys,collections
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
data=rows[2:]
base=int(data[0][0],16)
tot=sum(int(r[ix['Instructions Executed']]) for r in data)
tots=sum(int(r[ix['# Samples']]) for r in data)
print('total inst',tot,'samples',tots)
stalls=[h for h in hdr if h.startswith('stall_')]
agg=collections.Counter()
for r in data:
    for s_ in stalls: agg[s_]+=int(r[ix[s_]] or 0)
print([(k,round(100*v/tots,1)) for k,v in agg.most_common(8)])
blk=int(sys.argv[2],0) if len(sys.argv)>2 else 0x200
acc=collections.OrderedDict()
for r in data:
    a=int(r[0],16)-base
    d=acc.setdefault(a//blk,[0,0]); d[0]+=int(r[ix['Instructions Executed']]); d[1]+=int(r[ix['# Samples']])
for k,(i,s) in acc.items():
    if 100*i/tot>0.8 or 100*s/tots>0.8: print(hex(k*blk), f"inst {100*i/tot:5.1f}%  samples {100*s/tots:5.1f}%")
# top sampled instructions
top=sorted(data,key=lambda r:-int(r[ix['# Samples']]))[:14]
for r in top: print(hex(int(r[0],16)-base), r[1].strip()[:60], r[ix['# Samples']], round(100*int(r[ix['# Samples']])/tots,1))
