"""K1 fill time against the query length (random ACGT pairs, T = 150): does a trimmed last block pay?"""
import sys, json, ctypes as C
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
from bioinfo1_b200 import capi
ctx=capi.Context(0); L=capi.lib(); dev=torch.device('cuda',0)
n=1<<20; T=150
rng=np.random.default_rng(5)
for Q in [int(a) for a in sys.argv[1:]] or [128,136,144,150,152,160]:
    qb=np.frombuffer(b"ACGT",dtype=np.uint8)[rng.integers(0,4,size=n*Q)]; tb=np.frombuffer(b"ACGT",dtype=np.uint8)[rng.integers(0,4,size=n*T)]
    qo=(np.arange(n+1,dtype=np.uint64)*Q); to=(np.arange(n+1,dtype=np.uint64)*T)
    d_q=torch.from_numpy(qb).to(dev); d_t=torch.from_numpy(tb).to(dev)
    plan=C.c_void_p(); capi.check(L.b200_align_plan_create(ctx.h,n,qo.ctypes.data,to.ctypes.data,0,1,-1,-1,1,C.byref(plan)))
    cells=int(L.b200_align_plan_cells(plan)); cap=600*n+(1<<20)
    d_s=torch.empty(n,dtype=torch.int32,device=dev); d_b=torch.empty(n,dtype=torch.int32,device=dev)
    d_c=torch.empty(cap,dtype=torch.uint8,device=dev); d_o=torch.empty(n+1,dtype=torch.int64,device=dev)
    st=torch.cuda.current_stream()
    def step(): capi.check(L.b200_align_plan_run(plan,d_q.data_ptr(),d_t.data_ptr(),d_s.data_ptr(),d_b.data_ptr(),d_c.data_ptr(),d_o.data_ptr(),cap,st.cuda_stream))
    for _ in range(3): step()
    torch.cuda.synchronize()
    ctx.set_option("profile",1)
    fm=[]
    for _ in range(3):
        ctx.set_option("reset_counters",1); step(); torch.cuda.synchronize(); fm.append(ctx.counter("fill_ns")/1e6)
    ctx.set_option("profile",0)
    print(json.dumps({"Q":Q,"fill_ms":min(fm),"fill_gcups":cells/min(fm)/1e6,"ms_per_row":min(fm)/Q}))
    L.b200_align_plan_destroy(plan)
