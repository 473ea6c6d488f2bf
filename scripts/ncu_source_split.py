import csv,gzip,sys,pickle
rows=csv.reader(gzip.open('gpurun_out/mlp_src.csv.gz','rt'))
res=[];cur=None
for r in rows:
    if r and r[0]=='Kernel Name':
        cur={'name':r[1],'rows':[]}; res.append(cur); continue
    if cur is None or not r or r[0]=='Address': 
        if cur is not None and r and r[0]=='Address': cur['hdr']=r
        continue
    cur['rows'].append(r)
print(len(res))
pickle.dump(res,open('/tmp/mlp_src.pkl','wb'))
for i,k in enumerate(res): print(i,k['name'][:50],len(k['rows']), sum(int(r[5]) for r in k['rows']), sum(int(r[4]) for r in k['rows']))
