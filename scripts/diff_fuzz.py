"""differential fuzz on the CPU: the product's contractors (harness, compiled watch records incl. NE / LITS / LINEAR /
memo interpreter) against the oracle on random walks over generated models"""
import sys, os, random, ctypes as C
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
import numpy as np
import csolve_b200 as cb, util, gen_random
os.environ["NC"]="8"
hc=util.harness_lib()
seed0=int(sys.argv[1]); n=int(sys.argv[2])
models=checked=mism=lin=lov=rel=0
for s in range(seed0, seed0+n):
    text=gen_random.gen_instance(s)
    try: m=cb.Model(text)
    except cb.CsolveError: continue
    if hc.hc_load(C.byref(m.flat),1)!=0: continue
    models+=1; lin+= hc.hc_n_linear()>0; rel+= hc.hc_n_linrel()>0
    orc=util.Oracle(m); V=m.n_vars; rng=random.Random(s)
    for _w in range(12):
        dom=m.root_domains.copy()
        best = 2**31-1 if m.objective==cb.OBJ_MIN else -2**31
        free=[v for v in range(V) if v!=m.obj_var]; rng.shuffle(free)
        for v in free:
            lo,hi=int(dom[2*v]),int(dom[2*v+1])
            if hi-lo>64: val=rng.choice([lo,hi,lo+1,hi-1])
            else: val=rng.randint(lo,hi)
            if m.obj_var>=0 and rng.random()<0.3: best=rng.randint(-40,40)
            eo,ef=orc.node(dom,v,val,best)
            out=np.empty_like(dom)
            hf=hc.hc_node(util.p32(np.ascontiguousarray(dom,np.int32)),v,val,best,util.p32(out))
            checked+=1
            if not ef and (eo[0::2] > eo[1::2]).any():
                break        # documented deviation: an <obj> interval emptied by the incumbent fails the node at once
            if bool(hf)!=bool(ef) or (not ef and not np.array_equal(out,eo)):
                mism+=1
                if mism<=3: print("MISMATCH seed",s,"var",v,"val",val,"best",best,"\n",text,dom.tolist(),"\noracle",ef,eo.tolist(),"\nharness",hf,out.tolist())
            if not ef:
                # the register-resident forms (lane owns variable), when the model is a pure NOT(EQ) network
                out2=np.empty_like(dom)
                lf=hc.hc_node_lov(util.p32(np.ascontiguousarray(dom,np.int32)),v,val,util.p32(out2))
                if lf>=0:
                    lov+=1
                    if lf!=0 or not np.array_equal(out2,eo):
                        mism+=1
                        if mism<=3: print("MISMATCH (lov) seed",s,"var",v,"val",val,"\n",text,dom.tolist(),"\noracle",eo.tolist(),"\nlov",lf,out2.tolist())
            if ef: break
            dom=eo
print("seeds %d..%d: models %d (with linear clause %d, with small linear relations %d), node transitions %d, mismatches %d, of those also through the lane-owns-variable form %d"%(seed0,seed0+n,models,lin,rel,checked,mism,lov))
