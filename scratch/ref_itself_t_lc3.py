import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, time
from oracle import oracle as O, ref as R
from semantic_suma_b200 import synth
from helpers import sized, surfel_fields_equal
W=300; N=int(sys.argv[1]) if len(sys.argv)>1 else 130
p=O.default_params(**sized(W))
lp=dict(search_distance=3.0, min_trajectory_distance=15.0, min_verifications=2)
f=R.Full(p, **{'close-loops':True,'loop-search-distance':3.0,'loop-min-trajectory-distance':15.0,'loop-min-verifications':2})
scene=synth.Scene(width=W,height=64); poses=synth.trajectory(N, step=0.2618, yaw_deg=3.0)
O.gl_sums(1); osl=O.Slam(p); osl.enable_loop_closure(**lp)
for t in range(N):
    pts=scene.scan(t,poses[t])[0]
    if osl.loop_info()['optimisation_requested']:
        k=osl.integrate_loop_closures(); print('   integrated',k,'poses before scan',t)
    f.process_scan(pts); osl.process_scan(pts)
    info=osl.loop_info(); eq=np.array_equal(f.pose(),osl.pose())
    if t>=108 or not eq:
        same=True
        try: surfel_fields_equal(f.map_download(),osl.map.download())
        except AssertionError as e: same=str(e)[:100]
        print(t,'eq' if eq else 'DIFF %.1e'%np.abs(f.pose()-osl.pose()).max(), f.loop_flags(), (info['found_candidate'],info['use_candidate']), len(f.edges()), info['n_edges'], 'added',info['loop_edges_added'],'optreq',info['optimisation_requested'],'loop_count',info['loop_count'],'surfels',same)
