import csv,re,sys
src=open('/root/repo/mate_b200/csrc/mate_step.cuh').read().split('\n')
com=open('/root/repo/mate_b200/csrc/mate_common.cuh').read().split('\n')
def funcs(lines):
    out=[]
    for i,l in enumerate(lines):
        if l.startswith('__device__') or l.startswith('mate_step_kernel2'):
            name=re.findall(r'([A-Za-z_0-9]+)\(',l)
            if name: out.append((i+1,name[0]))
    return out
fs={'mate_step.cuh':funcs(src),'mate_common.cuh':funcs(com)}
rows=list(csv.reader(open(sys.argv[1])))
hdr=None; fname='?'; agg={}
for r in rows:
    if r and r[0]=='File Path': fname=r[1].split('/')[-1]; continue
    if r and r[0]=='Line No': hdr=r; ii=hdr.index('Instructions Executed'); isamp=hdr.index('# Samples'); ith=hdr.index('Thread Instructions Executed'); continue
    if hdr is None or len(r)<len(hdr): continue
    if r[0]!='' and r[2]=='-':
        ln=int(r[0]); name=fname
        if fname in fs:
            name='?'
            for s,n in fs[fname]:
                if ln>=s: name=n
        a=agg.setdefault(name,[0,0,0])
        try: a[0]+=int(r[ii]); a[1]+=int(r[isamp]); a[2]+=int(r[ith])
        except: pass
tot_s=sum(a[1] for a in agg.values())
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:16]:
    print('%-26s inst/env %6.1f  samples %5.1f%%  lanes %4.1f'%(k,a[0]/65536,100*a[1]/tot_s,a[2]/max(a[0],1)))
