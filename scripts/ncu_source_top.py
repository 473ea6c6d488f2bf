import pickle,sys
res=pickle.load(open('/tmp/mlp_src.pkl','rb'))
k=res[int(sys.argv[1])]
rows=k['rows']; base=int(rows[0][0],16)
tot=sum(int(r[5]) for r in rows); ts=sum(int(r[4]) for r in rows)
print(k['name'][:80], 'instr',tot,'samples',ts)
# classify by opcode
from collections import Counter
ci=Counter(); cs=Counter()
for r in rows:
    op=r[1].strip().split()
    o=op[1] if op[0].startswith('@') else op[0]
    o=o.split('.')[0]
    ci[o]+=int(r[5]); cs[o]+=int(r[4])
print('by opcode (instr share, sample share):')
for o,n in ci.most_common(28): print('  %-10s %5.1f%% %5.1f%%'%(o,100*n/tot,100*cs[o]/ts))
print('top sample lines:')
idx=sorted(range(len(rows)),key=lambda i:-int(rows[i][4]))[:int(sys.argv[2]) if len(sys.argv)>2 else 30]
for i in sorted(idx):
    r=rows[i]; print('  %05x n=%9s s=%5s  %s'%(int(r[0],16)-base,r[5],r[4],r[1].strip()[:70]))
