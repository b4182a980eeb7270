import sys, random
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, seqgen
from bioinfo1_b200 import capi
from cpu_checkers import load_oracle
O=load_oracle(); ctx=capi.Context(0)
rng=np.random.default_rng(5); pr=random.Random(5)
# long16 / long32 / generic small
for typ in (0,1,2):
    qs,ts=[],[]
    for n in (1,33,64,129,700,2100):
        t=seqgen.random_dna(rng,n); q=seqgen.mutate(rng,t,sub=0.03,ins=0.05,dele=0.05)
        qs.append(q.tobytes()); ts.append(t.tobytes())
    qs+= [b"", b"ACGTN-ACGT"*5]; ts+=[b"ACG", b"ACGTTACG-T"*4]
    got=ctx.align(qs,ts,typ)
    for q,t,g in zip(qs,ts,got): assert g==O.align(q,t,typ), (typ,len(q),len(t))
    ctx.set_option("long16",0); got=ctx.align(qs,ts,typ); ctx.set_option("long16",1)
    for q,t,g in zip(qs,ts,got): assert g==O.align(q,t,typ)
    # K1: 8192 small pairs
    qb,qo,tb,to=seqgen.short_pairs(3+typ,8192,length=(40,44,50)[typ])   # last row block trimmed to 8 / 16 / 24 rows
    s,b,c,o=ctx.align_packed(qb,qo,tb,to,typ)
    for k in range(0,8192,501):
        q=qb[int(qo[k]):int(qo[k+1])].tobytes(); t=tb[int(to[k]):int(to[k+1])].tobytes()
        assert (int(s[k]),int(b[k]),c[int(o[k]):int(o[k+1])].tobytes())==O.align(q,t,typ)
# minimizers
for k,w in ((15,5),(16,4),(3,9),(21,2)):
    seqs=[seqgen.random_dna(rng,int(rng.integers(k+w,900))).tobytes() for _ in range(12)]+[seqgen.random_dna(rng,4500).tobytes(), b"G"*200]
    got=ctx.minimize(seqs,k,w)
    for sq,g in zip(seqs,got):
        e=O.minimize(sq,k,w,True); assert all(np.array_equal(x,y) for x,y in zip(g,e))
# mapper
ref=seqgen.random_dna(rng,30000).tobytes()
reads=[seqgen.mutate(rng,np.frombuffer(ref[s:s+900],dtype=np.uint8),sub=0.03,ins=0.03,dele=0.03).tobytes() for s in (100,5000,20000)]
idx=capi.Index(ctx,ref,15,5,0.001); res,cg=idx.map_batch(reads,True,2,1,-1,-1,True); idx.close()
assert int(res["mapped"].sum())==3
print("sanitizer workload ok")
