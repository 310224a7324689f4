import csv,re,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
i0=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[i0]; data=[dict(zip(h,r)) for r in rows[i0+1:] if len(r)==len(h)]
tot=collections.Counter(); cnt=collections.Counter()
for d in data:
    n=re.sub(r"\(.*$","",d["Kernel Name"].replace("void ","")); n=re.sub(r"stg::<unnamed>::|at::native::|stg::","",n)[:64]
    tot[n]+=float(d["Metric Value"])/1e3; cnt[n]+=1
print(sys.argv[1], "total us", round(sum(tot.values())), "launches", len(data))
for n,v in tot.most_common(12): print(f"{v:8.1f} us x{cnt[n]:3d} {n}")
