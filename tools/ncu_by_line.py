import csv,sys,collections
rows=list(csv.reader(open(sys.argv[1])))
kern=sys.argv[2] if len(sys.argv)>2 else None
cur_file=None; cur_fn=None; hdr=None
agg=collections.OrderedDict()
tot=0; totS=0
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur_file=r[1].split('/')[-1]; continue
    if r[0]=="Function Name": cur_fn=r[1]; continue
    if r[0]=="Line No": hdr=r; iI=hdr.index("Instructions Executed"); iS=hdr.index("# Samples"); continue
    if kern and kern not in (cur_fn or ''): continue
    if r[0]!="":   # source line summary row
        key=(cur_file,int(r[0]),r[1].strip()[:90])
        a=agg.setdefault(key,[0,0]); vi=int(r[iI]) if r[iI].isdigit() else 0; vs=int(r[iS]) if r[iS].isdigit() else 0; a[0]+=vi; a[1]+=vs; tot+=vi; totS+=vs
print("total inst",tot,"samples",totS)
items=sorted(agg.items(), key=lambda kv:-kv[1][0])
for (f,ln,src),(i,s) in items[:int(sys.argv[3]) if len(sys.argv)>3 else 45]:
    print(f"{f:14s}:{ln:4d} inst {100*i/tot:5.1f}% samp {100*s/totS:5.1f}%  {src}")
