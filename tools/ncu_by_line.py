import csv,sys,collections
path=sys.argv[1]
rows=list(csv.reader(open(path)))
cur=None; cursrc=''
agg=collections.OrderedDict()
for r in rows:
    if len(r)>2 and r[0].isdigit():
        cur=int(r[0]); cursrc=r[1]
    elif len(r)>7 and r[2].startswith('0x'):
        try: s=int(r[6]); x=int(r[7])
        except: continue
        a=agg.setdefault(cur,[0,0,cursrc]); a[0]+=s; a[1]+=x
tots=sum(a[0] for a in agg.values()); totx=sum(a[1] for a in agg.values())
print('total samples',tots,'total warp instr',totx)
for ln,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 40]:
    print('%5s  samp %6.2f%%  instr %6.2f%%  %s'%(ln,100*a[0]/tots,100*a[1]/totx,a[2][:90]))
