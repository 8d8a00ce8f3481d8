import sys, time, json, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import fpt_b200.api as api, fpt_b200.synth as synth
from fpt_b200 import _lib
import ctypes as C
lib=_lib.load()
asize=bsize=500
nwin=int(sys.argv[1]) if len(sys.argv)>1 else 148
regend=50000*nwin
ch=synth.chromosome_fast(3, regend, 167*nwin, asize, bsize, wstep=50000)
lib.fpt_profile_enable(1)
lib.fpt_set_lanczos_form(int(__import__('os').environ.get('FPT_LANCZOS_FORM', '3')))
lib.fpt_set_k4_mode(int(__import__('os').environ.get('FPT_K4_MODE', '2')))
lib.fpt_set_lanczos_threads(int(__import__('os').environ.get('FPT_LANCZOS_THREADS', '256')))
for runs in (1000,):
    t=time.time()
    s,p,wr=api.css_scan(ch["acodes"],ch["bcodes"],ch["pos"],asize,bsize,regend,50000,50000,1000,runs,mds=0,seed=1)
    dt=time.time()-t
    buf=C.create_string_buffer(8192); lib.fpt_profile_summary(buf,8192)
    print(runs, 'windows', nwin, 'scored', int((wr==1).sum()) if wr is not None else None, 'sec', round(dt,2), buf.value.decode())
    print(s[:4], p[:4])
    ph=(C.c_ulonglong*8)()
    if lib.fpt_debug_umma_phases(ph)==0:
        tot=sum(ph) or 1
        print('umma phase share (dist, observed, shuffle, rows, contraction, decide, sweep):', [round(x/tot,3) for x in ph][:7], 'Mcycles per window', round(tot/1e6/max(1,nwin),2))
    k4=(C.c_ulonglong*4)()
    if lib.fpt_debug_k4_phases(k4)==0:
        tot=sum(k4) or 1
        print('k4 (tcgen05 genotype GEMM) phase share (expand, wait for MMAs, drain+store):', [round(x/tot,3) for x in k4][:3], 'Mcycles per window', round(tot/1e6/max(1,nwin),3))
    if lib.fpt_debug_lanczos_phases(ph)==0:
        steps=ph[7]; ph[7]=0
        tot=sum(ph) or 1
        print('lanczos steps per window', round(steps/max(1,nwin),1))
        print('lanczos phase share (-, means+codes, products, gram-schmidt, tri-solves, norms, coordinates):', [round(x/tot,3) for x in ph][:7], 'Mcycles per window', round(tot/1e6/max(1,nwin),2))
