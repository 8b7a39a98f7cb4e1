import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
top=int(sys.argv[2]) if len(sys.argv)>2 else 25
hdr=None; fname='?'; cur=None; out=[]
tot=0
for r in rows:
    if r and r[0]=='File Path': fname=r[1].split('/')[-1]; continue
    if r and r[0]=='Line No': hdr=r; isamp=hdr.index('# Samples'); ii=hdr.index('Instructions Executed'); idx={k:hdr.index(k) for k in ['stall_long_sb','stall_short_sb','stall_wait','stall_branch_resolving','stall_no_inst','stall_lg','stall_membar','stall_math','stall_mio']}; continue
    if hdr is None or len(r)<len(hdr): continue
    if r[0]!='' and r[2]=='-': cur=(fname,int(r[0])); continue
    if r[0]=='' and cur:
        try: n=int(r[isamp]); i=int(r[ii])
        except: continue
        tot+=n
        st=' '.join('%s=%s'%(k[6:],r[j]) for k,j in idx.items() if r[j] not in ('0','-'))
        out.append((n,i/65536,cur[0][:11],cur[1],r[3].strip()[:60],st))
out.sort(reverse=True)
print(tot)
for n,i,f,ln,s,st in out[:top]: print(n,'%.1f'%i,f,ln,s,'|',st)
