import os, sys, json, torch
sys.path.insert(0, os.getcwd())
from nano_hevc_b200 import _lib
L=_lib.lib(); dev=torch.device("cuda:0"); st=torch.cuda.current_stream().cuda_stream
for n in (16,32):
    B=(1<<20) if n==16 else (1<<18)
    rot=3
    xs=[torch.randint(-255,256,(B,n,n),device=dev,dtype=torch.int16) for _ in range(rot)]
    co=[torch.empty((B,n,n),dtype=torch.int32,device=dev) for _ in range(rot)]
    rs=[torch.empty((B,n,n),dtype=torch.int32,device=dev) for _ in range(rot)]
    def fwd():
        for x,c in zip(xs,co): L.nh_forward_transform(x.data_ptr(),0,c.data_ptr(),B,n,0,st)
    def inv():
        for c,r in zip(co,rs): L.nh_inverse_transform(c.data_ptr(),r.data_ptr(),B,n,0,st)
    for name,fn,bpp in (("fwd",fwd,6),("inv",inv,8)):
        fn(); torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/5/rot
        print(n,name,os.environ.get("NH_XF_OCC","4"),round(ms*1e3,1),"us",round(B*n*n*bpp/ms/1e6/6455.9,3))
