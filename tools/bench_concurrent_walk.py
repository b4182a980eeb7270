import sys, json, ctypes as C
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch, seqgen
from bioinfo1_b200 import capi
ctx=capi.Context(0); L=capi.lib(); dev=torch.device('cuda',0)
bq,bt=seqgen.ont_like_pairs(4242,256,mean_len=8000)
qs,ts=bq*8,bt*8
qb,qo=seqgen.pack_arrays(qs); tb,to=seqgen.pack_arrays(ts); n=len(qs)
d_q=torch.from_numpy(qb).to(dev); d_t=torch.from_numpy(tb).to(dev)
for cw in (1,0):
    ctx.set_option("concurrent_walk",cw)
    plan=C.c_void_p(); capi.check(L.b200_align_plan_create(ctx.h,n,qo.ctypes.data,to.ctypes.data,2,1,-1,-1,1,C.byref(plan)))
    cells=int(L.b200_align_plan_cells(plan)); cap=int(L.b200_align_plan_cigar_bound(plan))
    d_s=torch.empty(n,dtype=torch.int32,device=dev); d_b=torch.empty(n,dtype=torch.int32,device=dev)
    d_c=torch.empty(cap,dtype=torch.uint8,device=dev); d_o=torch.empty(n+1,dtype=torch.int64,device=dev)
    st=torch.cuda.current_stream()
    def step(): capi.check(L.b200_align_plan_run(plan,d_q.data_ptr(),d_t.data_ptr(),d_s.data_ptr(),d_b.data_ptr(),d_c.data_ptr(),d_o.data_ptr(),cap,st.cuda_stream))
    for _ in range(2): step()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(3): step()
    e1.record(st); torch.cuda.synchronize()
    print(json.dumps({"concurrent_walk":cw,"ms":e0.elapsed_time(e1)/3,"gcups":cells/(e0.elapsed_time(e1)/3)/1e6}))
    L.b200_align_plan_destroy(plan)
