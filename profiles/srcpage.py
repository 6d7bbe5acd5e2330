"""Reads `ncu -i REPORT --page source --csv --print-source sass --kernel-name regex:NAME` (one section per
launch) and prints the stall-reason shares, the executed warp instructions and the SASS lines with the
most stall samples.  Usage: python profiles/srcpage.py page.csv [section] [top_n].  Used for the
source-level findings quoted in r02_ncu_summary.md (uf_edges unions, walk round chain)."""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
secs=[];cur=None
for r in rows:
    if r and r[0]=='Kernel Name': cur={'name':r[1],'hdr':None,'data':[]}; secs.append(cur); continue
    if cur is None: continue
    if cur['hdr'] is None: cur['hdr']=r; continue
    cur['data'].append(r)
which=int(sys.argv[2]) if len(sys.argv)>2 else 0
ntop=int(sys.argv[3]) if len(sys.argv)>3 else 40
for k,s in enumerate(secs):
    ix={h:i for i,h in enumerate(s['hdr'])}
    tot=sum(int(r[ix['# Samples']]) for r in s['data'])
    print(k,s['name'][:60],'samples',tot,'instrs',len(s['data']))
s=secs[which]; hdr=s['hdr']; data=s['data']; ix={h:i for i,h in enumerate(hdr)}
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg={x:0 for x in stalls}
for r in data:
    for x in stalls: agg[x]+=int(r[ix[x]])
tot=sum(agg.values())
print([(k,round(100*v/tot,1)) for k,v in sorted(agg.items(), key=lambda x:-x[1])[:9]])
execd=sum(int(r[ix['Instructions Executed']]) for r in data)
print('warp instr executed',execd)
top=sorted(enumerate(data),key=lambda r:-int(r[1][ix['# Samples']]))[:ntop]
for i,r in top:
    st=sorted([(int(r[ix[x]]),x[6:]) for x in stalls],reverse=True)[:2]
    print(i, r[ix['# Samples']], r[ix['Instructions Executed']], r[ix['Source']].strip()[:64], st)
