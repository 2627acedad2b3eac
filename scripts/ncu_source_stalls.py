import csv, sys, re
fn = sys.argv[1]; kern = sys.argv[2] if len(sys.argv)>2 else None; topn = int(sys.argv[3]) if len(sys.argv)>3 else 45
rows=[]; cur=None; hdr=None; seen=0
for r in csv.reader(open(fn)):
    if r and r[0]=="Kernel Name":
        cur=r[1]; seen+=1; hdr=None; continue
    if r and r[0]=="Address": hdr=r; continue
    if hdr and cur and (kern is None or kern in cur) :
        rows.append((seen,dict(zip(hdr,r))))
first = rows[0][0]
rows=[d for s,d in rows if s==first]
tot=sum(int(d["# Samples"] or 0) for d in rows)
tinst=sum(int(d["Instructions Executed"] or 0) for d in rows)
print("total samples",tot,"total warp-instr",tinst, "n sass", len(rows))
stalls=[k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
agg={k:sum(int(d[k] or 0) for d in rows) for k in stalls}
print({k:v for k,v in sorted(agg.items(), key=lambda x:-x[1]) if v})
idx=sorted(range(len(rows)), key=lambda i:-int(rows[i]["# Samples"] or 0))[:topn]
for i in sorted(idx):
    d=rows[i]
    st={k[6:]:int(d[k]) for k in stalls if int(d[k] or 0)}
    top=sorted(st.items(), key=lambda x:-x[1])[:3]
    print(f'{i:5d} {int(d["# Samples"]):6d} {int(d["Instructions Executed"]):9d}  {d["Source"][:70]:70s} {top}')
