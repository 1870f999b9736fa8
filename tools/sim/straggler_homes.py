"""Development: where the queue entries of a marching row could go instead (bench map, CPU only).

For every broken parked sum (south-west `broke`, south-east `ebroke`) of the product scheme (east carry + vertical merge):
is its target a tap of this lane or of a lane one or two to the side in the SAME row?  Prints entries per 32-pixel row.
Result on the bench map: broke 3.0 per row, of which 0.79 land on the left neighbour's north-east tap, 0.67 on the right
neighbour's north-west, 0.56 on the left neighbour's north-west; ebroke 1.04, of which 0.37 land on the lane's OWN north-west
tap and 0.32 on the right neighbour's north-east: 2.7 of a row's 8.8 entries have a home one shuffle away."""
import sys, numpy as np, collections
import os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'tests'))
import scatter_sim as S
x0,y0=S.taps(); W=S.W
o_all=y0*W+x0
cnt=collections.Counter(); rows=0; nb=0; ne=0
for sr in range(128,128+256,4):
    for c0 in range(0,W-31,32):
        co=np.full(32,-1,np.int64); eo=np.full(32,-1,np.int64)
        for r in range(sr,sr+4):
            o=o_all[r,c0:c0+32]
            take=np.zeros(32,bool); take[1:]=o[:-1]+1==o[1:]
            given=np.zeros(32,bool); given[:-1]=take[1:]
            chain=co==o; vd=co==o+W
            broke=(co>=0)&~chain&~vd
            e_chain=eo==o+1; e_vd=eo==o+W+1
            e_broke=(eo>=0)&~e_chain&~e_vd
            rows+=1
            for l in np.nonzero(broke)[0]:
                nb+=1; t=co[l]; hit=None
                for d in (-1,1,-2,2):
                    k=l+d
                    if 0<=k<32:
                        for name,off in (("NW",0),("NE",1),("SW",W),("SE",W+1)):
                            if t==o[k]+off: hit=(d,name); break
                    if hit: break
                cnt[("broke",hit)]+=1
            for l in np.nonzero(e_broke)[0]:
                ne+=1; t=eo[l]; hit=None
                for d in (0,-1,1,2):
                    k=l+d
                    if 0<=k<32:
                        for name,off in (("NW",0),("NE",1),("SW",W),("SE",W+1)):
                            if t==o[k]+off: hit=(d,name); break
                    if hit: break
                cnt[("ebroke",hit)]+=1
            co=o+W; eo=np.where(~given,o+W+1,-1)
for k,v in cnt.most_common(20): print(k, round(v/rows,3))
print("broke/row",nb/rows,"ebroke/row",ne/rows)
