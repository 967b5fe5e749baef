#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 -k "fused or gate or block or decode or spectre_base or goldens or sweep or long_context or race" 2>&1 | tail -8 > gpurun_out/r02c_pytest_gpu.log
tail -8 gpurun_out/r02c_pytest_gpu.log
for sched in 3 19; do
AB_SCHED=$sched timeout 300 python tools/ab_anchors.py >> gpurun_out/r02c_ab_anchors.log 2>&1; tail -1 gpurun_out/r02c_ab_anchors.log
done
AB_BATCH=32 AB_NFFT=1024 timeout 300 python tools/ab_anchors.py >> gpurun_out/r02c_ab_anchors.log 2>&1; tail -1 gpurun_out/r02c_ab_anchors.log
AB_SCHED=19 AB_BATCH=8 timeout 300 python tools/ab_anchors.py >> gpurun_out/r02c_ab_anchors.log 2>&1; tail -1 gpurun_out/r02c_ab_anchors.log
AB_BATCH=148 timeout 300 python tools/ab.py -350,3,0 -350,19,0 > gpurun_out/r02c_ab_pair.log 2>&1; cat gpurun_out/r02c_ab_pair.log
AB_NFFT=16384 AB_BATCH=16 timeout 300 python tools/ab.py -350,3,0 -350,19,0 > gpurun_out/r02c_ab_pair_16384.log 2>&1; cat gpurun_out/r02c_ab_pair_16384.log
# parity of the paired order against the default order (bit-equal expected)
timeout 300 python - <<'PY' 2>&1 | tail -5
import torch, fft_b200
from fft_b200 import _lib
lib=_lib.load()
torch.manual_seed(0)
for (B,N,C,dt) in [(7,4096,768,torch.float32),(3,4000,64,torch.float32),(5,4096,96,torch.bfloat16),(2,16384,64,torch.float32)]:
    V=torch.randn(B,N,C,device='cuda').to(dt); g=torch.randn(B,C//16,(4096 if N<=4096 else 16384)//2+1,dtype=torch.cfloat,device='cuda')
    nf=4096 if N<=4096 else 16384
    lib.spectre_mix_set_sched(3); a=fft_b200.spectral_mix(V,g,n_fft=nf,group_width=16)
    lib.spectre_mix_set_sched(19); b=fft_b200.spectral_mix(V,g,n_fft=nf,group_width=16)
    lib.spectre_mix_set_sched(3)
    print((B,N,C,str(dt)), 'paired == default:', torch.equal(a,b), float((a.float()-b.float()).abs().max()))
PY
