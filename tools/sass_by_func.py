import csv,re,sys,bisect
srccsv, dis, iters = sys.argv[1], sys.argv[2], float(sys.argv[3])
rows=list(csv.reader(open(srccsv)))
hdr=rows[1]; ix={n:i for i,n in enumerate(hdr)}
inst=[r for r in rows[2:] if len(r)==len(hdr)]
lines=open(dis).read().split("\n")
kernel='_ZN2rb11scan_kernelILi32EEEvNS_8DevBatchE'
start=next(i for i,l in enumerate(lines) if l.startswith(".text."+kernel+":"))
loc=None; locs=[]
for l in lines[start+1:]:
    if l.startswith("//----") and ".text." in l: break
    m=re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        loc=[(m.group(1).split('/')[-1], int(m.group(2)))]+[(a.split('/')[-1],int(b)) for a,b in re.findall(r'inlined at "([^"]+)", line (\d+)', l)]
        continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", l): locs.append(loc)
src=open('/root/repo/ribbit_b200/csrc/scan_core.h').read().split('\n')
funcs=[]
for i,l in enumerate(src,1):
    m=re.match(r'^RB_HD \S+ (\w+)\(', l) or re.match(r'^RB_HD \S+ \S+ (\w+)\(',l)
    if m: funcs.append((i,m.group(1)))
def fn(line):
    k=bisect.bisect_right([f[0] for f in funcs], line)-1
    return funcs[k][1] if k>=0 else '?'
agg={}; tot=0
for k in range(min(len(inst), len(locs))):
    ie=int(inst[k][ix["Instructions Executed"]]); te=int(inst[k][ix["Thread Instructions Executed"]])
    ch=locs[k]; key='?'
    if ch:
        sc=[c for c in ch if c[0]=='scan_core.h']
        if sc:
            names=[fn(c[1]) for c in sc]
            key=names[-1]
            for nme in reversed(names):
                if nme not in ('lane_phase2','lane_phase1'): key=nme; break
        else:
            kc=[c for c in ch if c[0]=='kernels.cu']
            key='kernels.cu:%d'%kc[-1][1] if kc else ch[-1][0]
    a=agg.setdefault(key,[0,0]); a[0]+=ie; a[1]+=te; tot+=ie
print('total warp instr %.3g = %.1f per iteration' % (tot, tot/iters))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:45]:
    print('%-22s %6.2f%%  %6.1f instr/iter  active %.1f' % (k, 100*v[0]/tot, v[0]/iters, v[1]/max(v[0],1)))
