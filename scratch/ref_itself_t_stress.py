import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, time
from oracle import oracle as O, ref as R
from semantic_suma_b200 import synth
from helpers import sized, surfel_fields_equal
O.gl_sums(1)
W=900
for seed,step,yaw,sem in ((7,1.0,0.5,True),(11,0.6,2.0,True),(23,1.4,-1.0,False),(31,0.3,4.0,True)):
    p=O.default_params(**sized(W))
    scene=synth.Scene(width=W,height=64,seed=seed,semantic=sem)
    N=30
    poses=synth.trajectory(N, step=step, yaw_deg=yaw)
    f=R.Full(p); osl=O.Slam(p); res='ok'
    t0=time.time()
    for t in range(N):
        sc=scene.scan(t,poses[t])
        f.process_scan(*sc); osl.process_scan(*sc)
        if not np.array_equal(f.pose(),osl.pose()): res='t=%d pose differs %.2e'%(t,np.abs(f.pose()-osl.pose()).max()); break
        if f.map_size()!=osl.map.size(): res='t=%d size %d vs %d'%(t,f.map_size(),osl.map.size()); break
        if t%5==4 or t==N-1:
            try: surfel_fields_equal(f.map_download(),osl.map.download())
            except AssertionError as e: res='t=%d %s'%(t,str(e)[:120]); break
    print(seed,step,yaw,sem,res,'track_loss',osl.stats()['track_loss'],'surfels',osl.map.size(),'%.0fs'%(time.time()-t0),flush=True)
