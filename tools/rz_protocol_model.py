#!/usr/bin/env python
"""Randomised model of k_resize_strips' warp-role hand-over (producer / 16 horizontal / 6 vertical warps, mbarrier phases and
parities as in csrc/resize_strips.cu): every interleaving must finish, no horizontal warp may write chunk j before the vertical
pass of chunk j - 2 is done, no vertical warp may start chunk j before all 16 horizontal warps arrived, a sub-stage slot is only
refilled after its 8 readers released it.  Run: python tools/rz_protocol_model.py"""
import random
class MBar:
    def __init__(s,count): s.count=count; s.pending=count; s.phase=0
    def arrive(s):
        s.pending-=1
        assert s.pending>=0
        if s.pending==0: s.phase+=1; s.pending=s.count
    def test(s,parity): return (s.phase&1)!=parity   # phase with given parity completed
def run(ns,nchunks,HW=16,VW=6,seed=0):
    rnd=random.Random(seed)
    SUBS=4; HG=2
    full=[MBar(1) for _ in range(ns)]; empty=[MBar(8) for _ in range(ns)]
    hdone=[MBar(HW),MBar(HW)]; vdone=[MBar(VW),MBar(VW)]
    slot_content=[None]*ns   # (chunk,sub)
    log={'hwrite':{}, 'vdone':set(), 'hdone_cnt':{}}
    def producer():
        q=0;par=0;round0=True
        for k in range(nchunks):
            for sub in range(SUBS):
                if not round0:
                    while not empty[q].test(par^1): yield
                slot_content[q]=(k,sub)
                full[q].arrive()
                yield
                q+=1
                if q==ns: q=0;par^=1;round0=False
    def hwarp(w):
        hgrp=w//8
        qc=0;parc=0
        for j in range(nchunks):
            qs=[]
            for i in range(2):
                a=qc+hgrp+i*HG;p=parc
                while a>=ns: a-=ns;p^=1
                qs.append((a,p))
            while not full[qs[0][0]].test(qs[0][1]): yield
            if j>=2:
                while not vdone[j&1].test(((j-2)>>1)&1): yield
                assert (j-2) in log['vdone'], ("H before V", j)
            for i in range(2):
                a,p=qs[i]
                if i>0:
                    while not full[a].test(p): yield
                assert slot_content[a]==(j,hgrp+i*HG),(slot_content[a],j,hgrp,i)
                yield
                # write ring rows of chunk j: V(j-2) must be done, V(j-1)... (allowed)
                empty[a].arrive()
                yield
            log['hdone_cnt'][j]=log['hdone_cnt'].get(j,0)+1
            hdone[j&1].arrive()
            yield
            qc+=SUBS
            while qc>=ns: qc-=ns;parc^=1
    vcount={}
    def vwarp(v):
        for j in range(nchunks):
            while not hdone[j&1].test((j>>1)&1): yield
            assert log['hdone_cnt'].get(j,0)==HW,("V before H",j,log['hdone_cnt'].get(j))
            # H must not have started writing chunk j+2
            yield
            vcount[j]=vcount.get(j,0)+1
            if vcount[j]==VW: log['vdone'].add(j)
            vdone[j&1].arrive()
            yield
    threads=[producer()]+[hwarp(w) for w in range(HW)]+[vwarp(v) for v in range(VW)]
    alive=list(range(len(threads)))
    steps=0
    while alive:
        i=rnd.choice(alive)
        # bias: sometimes starve some
        try: next(threads[i])
        except StopIteration: alive.remove(i)
        steps+=1
        if steps>5_000_000: raise RuntimeError("deadlock? ns=%d"%ns)
    return True
for ns in range(2,9):
    for n in (1,2,3,7,20):
        for seed in range(3):
            run(ns,n,seed=seed)
print("ok")
