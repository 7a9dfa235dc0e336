import sys, ctypes as C
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from barkit_b200 import _lib
from oracle import bksyn
P=int(sys.argv[1]) if len(sys.argv)>1 else 2000000
prog=_lib.Program("^(?P<CB>[ATGCN]{16})atgccat(?P<UMI>[ATGCN]{12})",1)
ctx=_lib.Context(prog,None,paired=True)
cfg=_lib.syn_config(150,plant1=(b"ATGCCAT",16,0)); rl=329
ms=[]
for mate in (1,2):
    dt=ctx.dev_alloc(P*rl+64); do=ctx.dev_alloc((P+1)*4)
    ctx.generate_device(cfg,mate,20240902,0,P,dt,do)
    m=_lib.MateIn(); m.text,m.rec_off,m.text_len,m.max_rec_bytes=dt,do,P*rl,rl; ms.append(m)
cap=P*(rl+80); o1=ctx.dev_alloc(cap); o2=ctx.dev_alloc(cap)
for i in range(3): r=ctx.extract_device(P,ms[0],ms[1],o1,cap,o2,cap)
out=np.zeros(24,np.uint64)
_lib.lib.bk_debug_phases(ctx.handle,out.ctypes.data)
ntiles=(P+31)//32
print("kernel_ms",r.kernel_ms,"Mpairs/s",P/r.kernel_ms/1e3,"tiles",ntiles)
names={0:"L wait_empty",1:"L claim",2:"L offsets+issue",
 4:"S wait_pub",5:"S lookback (per tile of this scan warp x0.5)",
 8:"F wait_full",9:"F parse",10:"F match",11:"F exchange",12:"F size/scan/pub/meta",
 16:"B emit m0 part0",17:"B emit m0 part1",18:"B emit m0 part2",19:"B emit m0 part3",20:"B emit m1 q0",21:"B emit m1 q1",22:"B emit m1 q2",23:"B emit m1 q3"}
for i,n in names.items(): print("%-24s %8.0f cyc/tile"%(n,out[i]/ntiles))
