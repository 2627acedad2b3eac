import csv, sys
fn=sys.argv[1]; kern=sys.argv[2]
rows=[]; cur=None; hdr=None; seen=0
for r in csv.reader(open(fn)):
    if r and r[0]=="Kernel Name": cur=r[1]; seen+=1; hdr=None; continue
    if r and r[0]=="Address": hdr=r; continue
    if hdr and cur and kern in cur: rows.append((seen,dict(zip(hdr,r))))
first=rows[0][0]; rows=[d for s,d in rows if s==first]
# segment by changes in executed count
seg=[]; 
for i,d in enumerate(rows):
    n=int(d["Instructions Executed"] or 0); s=int(d["# Samples"] or 0)
    op=d["Source"].split()[0] if not d["Source"].startswith("@") else d["Source"].split()[1]
    if seg and abs(seg[-1][2]-n)<=0.02*max(n,1) : seg[-1][1]=i; seg[-1][3]+=n; seg[-1][4]+=s; seg[-1][5][op.split('.')[0]]=seg[-1][5].get(op.split('.')[0],0)+1
    else: seg.append([i,i,n,n,s,{op.split('.')[0]:1}])
tot=sum(x[3] for x in seg)
for a,b,n,t,s,ops in seg:
    if t>0.004*tot:
        top=sorted(ops.items(), key=lambda x:-x[1])[:8]
        print(f"{a:5d}-{b:5d} len {b-a+1:4d} exec/instr {n:9d} total {t:10d} ({100*t/tot:4.1f}%) samples {s:5d}  {top}")
