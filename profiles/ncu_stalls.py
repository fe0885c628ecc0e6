import csv, re, sys
from collections import defaultdict
src_csv, sass, kern = sys.argv[1:4]
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith("//--------------------- .text.") and kern in l)
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith("//--------------------- ")), len(lines))
cur=("?",0); seq=[]
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur=(m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l): seq.append(cur)
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg=defaultdict(lambda: defaultdict(float)); tot=defaultdict(float)
for k,r in enumerate(data):
    for s in stalls:
        v=float(r[ix[s]] or 0); agg[seq[k]][s]+=v; tot[s]+=v
T=sum(tot.values())
print('overall stall mix:', {s:round(100*v/T,1) for s,v in sorted(tot.items(), key=lambda kv:-kv[1])[:8]})
for key,d in sorted(agg.items(), key=lambda kv:-sum(kv[1].values()))[:18]:
    tt=sum(d.values())
    top=sorted(d.items(), key=lambda kv:-kv[1])[:3]
    print('%5.1f%%  %s:%d  '%(100*tt/T,key[0],key[1]), [(s.replace('stall_',''), round(100*v/tt)) for s,v in top])
