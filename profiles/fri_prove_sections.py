import os, sys, time
ROOT="/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from stark_brainfuck_b200 import Engine, mirror
from stark_brainfuck_b200 import glue as G
from stark_brainfuck_b200.glue import DeviceCodeword, Glue
from util import root_of_unity
mirror.register(); eng=Engine(0); glue=Glue(mirror.binding, eng); mirror.set_glue(glue); m=mirror
logn=20; n=1<<logn
coeffs=np.random.default_rng(logn).integers(0,18446744069414584321,(3,n//4),dtype=np.uint64)
planes=eng.ntt(eng.upload(coeffs),logn,root_of_unity(logn),offset=7)
fri=m.fri.Fri(m.field.generator(),m.field.primitive_nth_root(n),n,4,8,m.xfield)
T={}
def wrap(obj,name,key):
    f=getattr(obj,name)
    def w(*a,**k):
        t0=time.perf_counter(); r=f(*a,**k); T[key]=T.get(key,0)+time.perf_counter()-t0; return r
    setattr(obj,name,w)
wrap(G,"prefetch","prefetch_leaves"); wrap(G,"prefetch_paths","prefetch_paths")
wrap(glue,"fri_commit","commit"); wrap(glue,"fri_query","query"); wrap(glue,"fri_query_last","query_last")
wrap(G.NodeView,"__getitem__","commit:nodes[k] (root download)"); wrap(eng,"fri_fold","commit:fri_fold call")
wrap(glue,"merkle_build","commit:merkle_build"); wrap(m.ip.ProofStream,"prover_fiat_shamir","fiat_shamir")
wrap(m.xfield.__class__,"sample","sample")
for it in range(4):
    T.clear(); ps=m.ip.ProofStream(); torch.cuda.synchronize(); t0=time.perf_counter()
    fri.prove(DeviceCodeword(glue,planes,m.xfield),ps); torch.cuda.synchronize(); tot=time.perf_counter()-t0
print("total %.2f ms"%(tot*1e3), {k:round(v*1e3,2) for k,v in T.items()})
