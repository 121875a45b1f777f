import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, time
from oracle import oracle as O, ref as R
from helpers import scans, sized, surfel_fields_equal
O.gl_sums(1)
variants=[dict(), dict(weighting=1), dict(weighting=2), dict(weighting=0), dict(bilinear_sampling=0), dict(compose_rendering=0), dict(initialize_identity=1), dict(initialize_identity=0),
 dict(update_always=1), dict(weighting_scheme=1), dict(weighting_scheme=2), dict(averaging_scheme=1), dict(confidence_mode=0), dict(confidence_mode=1), dict(confidence_mode=2),
 dict(use_stability=0), dict(unstable_age=1, confidence_threshold=5.0), dict(max_iterations=3), dict(fallback_mode=0), dict(active_timestamps=1), dict(min_radius=0.05,max_radius=0.2), dict(max_angle=60.0), dict(map_max_distance=0.05, map_max_angle=5.0), dict(partial_extraction=0), dict(submap_extent=3.0, submap_dimension=1)]
W=450
for sem in (False,True):
    sc,_=scans(W,n=4,semantic=sem)
    for kw in variants:
        p=O.default_params(**sized(W),**kw)
        try:
            f=R.Full(p); osl=O.Slam(p)
            res='ok'
            for t in range(4):
                f.process_scan(*sc[t]); osl.process_scan(*sc[t])
                if not np.array_equal(f.pose(),osl.pose()): res='t=%d pose differs %.2e'%(t,np.abs(f.pose()-osl.pose()).max()); break
                if f.map_size()!=osl.map.size(): res='t=%d size %d vs %d'%(t,f.map_size(),osl.map.size()); break
                try: surfel_fields_equal(f.map_download(),osl.map.download())
                except AssertionError as e: res='t=%d %s'%(t,str(e)[:120]); break
        except Exception as e:
            res='EXC '+str(e)[:150]
        print('sem' if sem else 'geo',kw,res, flush=True)
