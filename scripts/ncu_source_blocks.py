"""Summarise an ncu SASS source page (scripts/ncu_source.sh): runs of instructions with the same execution count, their share
of the executed warp instructions and their stall samples.  python scripts/ncu_source_blocks.py file.csv.gz [min share]"""
import csv,gzip,sys
rows=list(csv.reader(gzip.open(sys.argv[1],'rt')))[2:]
base=int(rows[0][0],16)
tot=sum(int(r[5]) for r in rows); print('total instr',tot, 'samples', sum(int(r[4]) for r in rows))
blocks=[];cur=None
for r in rows:
    a=(int(r[0],16)-base); n=int(r[5]); s=int(r[4])
    if cur and abs(n-cur['n'])<=0.02*max(n,cur['n'],1):
        cur['end']=a; cur['sum']+=n; cur['cnt']+=1; cur['samp']+=s
    else:
        if cur: blocks.append(cur)
        cur=dict(start=a,end=a,n=n,sum=n,cnt=1,samp=s,first=r[1].strip())
blocks.append(cur)
thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.004
for b in blocks:
    if b['sum']>thr*tot:
        print('%05x-%05x  n=%9d  x%3d  share=%5.1f%%  samples=%5d  %s'%(b['start'],b['end'],b['n'],b['cnt'],100*b['sum']/tot,b['samp'],b['first'][:50]))
