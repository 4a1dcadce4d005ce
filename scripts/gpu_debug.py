import sys, os, ctypes as C, subprocess, numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from genarchbench_b200 import pairio, bsw
import oracle
d=os.path.join(ROOT,'tests','host_emul'); so=os.path.join(d,'libbsw_emul.so')
subprocess.run(['/usr/bin/g++','-O2','-std=c++17','-fPIC','-fopenmp','-shared','-w',f'-I{d}',f'-I{ROOT}/genarchbench_b200/csrc',f'-I{ROOT}/include','-o',so,os.path.join(d,'emul_lib.cpp')],check=True)
L=C.CDLL(so); L.bsw_emul_batch.argtypes=[C.c_void_p]*4+[C.c_int64,C.c_int32]
def has_n(b,k):
    p=b.pairs[k]; return bool((b.ref[p['idr']:p['idr']+p['len1']]==4).any() or (b.qer[p['idq']:p['idq']+p['len2']]==4).any())
def study(name,b,w,params=None):
    a=b.copy(); e=b.copy(); g=b.copy()
    oracle.oracle_batch(a,w=w,params=params)
    L.bsw_emul_batch(oracle._params_array(params), e.pairs.ctypes.data,e.ref.ctypes.data,e.qer.ctypes.data,len(e),w)
    with bsw.BswGpu(**(params or {})) as G:
        G.batch(g.pairs,g.ref,g.qer,w); st=G.stats()
    dg=(a.outputs()!=g.outputs()).any(axis=1); de=(a.outputs()!=e.outputs()).any(axis=1)
    bad=np.nonzero(dg)[0]
    print(f"== {name}: n={len(b)} gpu_bad={dg.sum()} emul_bad={de.sum()} short={st['pairs_short']} long={st['pairs_long']} launches={st['kernel_launches']}")
    if len(bad):
        wide=np.array([has_n(b,k) for k in bad])
        print("   bad wide frac", wide.mean(), " len2 range", b.pairs['len2'][bad].min(), b.pairs['len2'][bad].max(), "len1 range", b.pairs['len1'][bad].min(), b.pairs['len1'][bad].max(), "h0 range", b.pairs['h0'][bad].min(), b.pairs['h0'][bad].max())
        allwide=np.array([has_n(b,k) for k in range(min(len(b),3000))]).mean()
        print("   overall wide frac", allwide)
        for k in bad[:6]:
            print("    ",k,'len1',b.pairs['len1'][k],'len2',b.pairs['len2'][k],'h0',b.pairs['h0'][k],'N',has_n(b,k),'want',a.outputs()[k],'gpu',g.outputs()[k])
c=pairio.preset(4); c.len2_min,c.len2_max,c.h0_min,c.h0_max,c.n_frac,c.random_frac=1,300,0,80,0.3,0.2
for w in (1,3,10):
    study(f"band w={w}", pairio.generate(c,20000,seed=500+w), w)
c2=pairio.preset(4); c2.len2_min,c2.len2_max,c2.h0_min,c2.h0_max,c2.n_frac,c2.random_frac=1,300,0,80,0.0,0.2
study("band w=3 noN", pairio.generate(c2,20000,seed=503), 3)
c3=pairio.preset(1); c3.seed=7109; c3.sub_rate=0.1; c3.indel_rate=0.05
study("asym", pairio.generate(c3,1500), 100, dict(o_del=5,e_del=2,o_ins=7,e_ins=1))
study("gape2", pairio.generate(c3,1500), 100, dict(e_del=2,e_ins=2,zdrop=30))
study("default same data", pairio.generate(c3,1500), 100)
