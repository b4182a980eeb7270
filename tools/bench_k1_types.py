import sys, json, ctypes as C
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch, seqgen
from bioinfo1_b200 import capi
ctx=capi.Context(0); L=capi.lib(); dev=torch.device('cuda',0)
qb,qo,tb,to=seqgen.short_pairs(1000,1<<20)
d_q=torch.from_numpy(qb).to(dev); d_t=torch.from_numpy(tb).to(dev)
n=1<<20
for typ in (0,2,1):
    plan=C.c_void_p(); capi.check(L.b200_align_plan_create(ctx.h,n,qo.ctypes.data,to.ctypes.data,typ,1,-1,-1,1,C.byref(plan)))
    cells=int(L.b200_align_plan_cells(plan)); cap=64*n+(1<<20)
    d_s=torch.empty(n,dtype=torch.int32,device=dev); d_b=torch.empty(n,dtype=torch.int32,device=dev)
    d_c=torch.empty(cap,dtype=torch.uint8,device=dev); d_o=torch.empty(n+1,dtype=torch.int64,device=dev)
    st=torch.cuda.current_stream()
    def step(): capi.check(L.b200_align_plan_run(plan,d_q.data_ptr(),d_t.data_ptr(),d_s.data_ptr(),d_b.data_ptr(),d_c.data_ptr(),d_o.data_ptr(),cap,st.cuda_stream))
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10): step()
    e1.record(st); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    ctx.set_option("profile",1); ctx.set_option("reset_counters",1); step(); torch.cuda.synchronize()
    print(json.dumps({"type":typ,"ms":ms,"gcups":cells/ms/1e6,"fill_ms":ctx.counter("fill_ns")/1e6,"walk_ms":ctx.counter("walk_ns")/1e6}))
    ctx.set_option("profile",0)
    L.b200_align_plan_destroy(plan)
