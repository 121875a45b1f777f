"""unpinned build of the reference (nothing follows the oracle's rules) against the oracle in its default mode (exact sums = the
CUDA contract) over a longer run: pose difference and surfel-count difference per scan (no drift expected)"""
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np
from oracle import oracle as O, ref as R
from semantic_suma_b200 import synth
from helpers import sized
O.gl_sums(0)
W=900; N=40
p=O.default_params(**sized(W))
scene=synth.Scene(width=W,height=64,semantic=True); poses=synth.trajectory(N)
f=R.Full(p,mode="precise"); osl=O.Slam(p)
worst=0
for t in range(N):
    sc=scene.scan(t,poses[t]); f.process_scan(*sc); osl.process_scan(*sc)
    d=np.abs(f.pose()-osl.pose()).max(); worst=max(worst,d)
    if t%5==4: print(t,'pose diff %.2e'%d,'surfels',f.map_size(),osl.map.size(),'rel %.1e'%(abs(f.map_size()-osl.map.size())/osl.map.size()),flush=True)
print('worst pose difference over',N,'scans: %.2e m'%worst)
