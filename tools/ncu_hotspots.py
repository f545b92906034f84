import csv, subprocess, sys, io
rep, kern, out = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--kernel-name",f"regex:{kern}"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
hdr=rows[1]; isamp=hdr.index("# Samples"); isrc=hdr.index("Source"); iex=hdr.index("Instructions Executed")
stall=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data=[]
for r in rows[2:]:
    if len(r)!=len(hdr) or not r[isamp].isdigit():
        if data: break
        continue
    data.append(r)
tot=sum(int(r[isamp]) for r in data)
with open(out,'w') as f:
    f.write(f"# {rows[0][1][:80]}\n# total warp-stall samples {tot}; by reason: ")
    agg={hdr[i]:sum(int(r[i]) for r in data) for i in stall}
    f.write(", ".join(f"{k} {v}" for k,v in sorted(agg.items(), key=lambda kv:-kv[1])[:8])+"\n# sass_line samples executed instruction  top-2 stall reasons\n")
    top=sorted(range(len(data)), key=lambda i:-int(data[i][isamp]))[:30]
    for i in sorted(top):
        r=data[i]; st={hdr[c]:int(r[c]) for c in stall if int(r[c])>0}
        f.write(f"{i} {r[isamp]} {r[iex]} {r[isrc].strip()[:70]}  {sorted(st.items(), key=lambda kv:-kv[1])[:2]}\n")
print('wrote',out)
