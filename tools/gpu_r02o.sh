#!/bin/bash
mkdir -p gpurun_out
timeout 600 python - <<'PY' 2>&1 | tail -14
import torch, fft_b200, time, json
from fft_b200 import _lib
lib=_lib.load()
torch.manual_seed(0)
dev='cuda'
def run(sched, V, g, m=None):
    lib.spectre_mix_set_sched(sched)
    try:
        return fft_b200.spectral_mix(V, g, m, n_fft=8192, group_width=16)
    finally:
        lib.spectre_mix_set_sched(3)
for (B,N,C,mem) in [(2,8192,32,False),(3,8000,768,True),(1,8192,64,False),(2,5000,48,True)]:
    V=torch.randn(B,N,C,device=dev); g=torch.randn(B,C//16,4097,dtype=torch.cfloat,device=dev)
    m=torch.randn(4097,C,dtype=torch.cfloat,device=dev)/8 if mem else None
    a=run(3,V,g,m); b=run(35,V,g,m)
    Vf=torch.fft.rfft(V.double(),n=8192,dim=1); gb=g.to(torch.complex128).permute(0,2,1).repeat_interleave(16,-1)
    mix=gb*Vf + (m.to(torch.complex128) if mem else 0)
    want=torch.fft.irfft(mix,n=8192,dim=1)[:,:N].float()
    print((B,N,C,mem),'DIT2 err',float((a-want).norm()/want.norm()),'plain single-kernel err',float((b-want).norm()/want.norm()), fft_b200.plan_info(B,N,8192,C,16)['dit'])
B,C=32,768
V=[torch.randn(B,8192,C,device=dev) for _ in range(2)]; g=[torch.randn(B,C//16,4097,dtype=torch.cfloat,device=dev) for _ in range(2)]
alg=fft_b200.plan_info(B,8192,8192,C,16)['algorithmic_bytes']
for sched in (3,35,3,35):
    lib.spectre_mix_set_sched(sched)
    for i in range(3): fft_b200.spectral_mix(V[i%2],g[i%2],n_fft=8192,group_width=16)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(8): fft_b200.spectral_mix(V[i%2],g[i%2],n_fft=8192,group_width=16)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/8
    print(json.dumps({'sched':sched,'dit2':sched==3,'us':round(ms*1e3,1),'GBps':round(alg/ms/1e6)}))
    time.sleep(0.3)
lib.spectre_mix_set_sched(3)
PY
