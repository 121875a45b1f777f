import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, time
from oracle import oracle as O, ref as R
from semantic_suma_b200 import synth
from helpers import surfel_fields_equal
O.gl_sums(1)
cases=[dict(data_width=450,data_height=64,model_width=512,model_height=64),
       dict(data_width=450,data_height=64,model_width=450,model_height=96),
       dict(data_width=600,data_height=32,model_width=300,model_height=32),
       dict(data_width=450,data_height=64,model_width=450,model_height=64,data_fov_up=10.0,data_fov_down=-30.0,model_fov_up=10.0,model_fov_down=-30.0),
       dict(data_width=450,data_height=64,model_width=450,model_height=64,min_depth=1.0,max_depth=40.0,model_min_depth=1.0,model_max_depth=40.0),
       dict(data_width=450,data_height=64,model_width=450,model_height=64,model_fov_up=5.0,model_fov_down=-28.0),
       dict(data_width=450,data_height=64,model_width=450,model_height=64,icp_max_distance=0.5,icp_max_angle=10.0, factor=0.2),
       dict(data_width=450,data_height=64,model_width=450,model_height=64,p_stable=0.9,p_prior=0.6,sigma_angle=0.5,sigma_distance=0.05,max_weight=5.0),
      ]
for kw in cases:
    for sem in (False,True):
        p=O.default_params(**kw)
        scene=synth.Scene(width=kw['data_width'],height=kw['data_height'],fov_up=kw.get('data_fov_up',3.0),fov_down=kw.get('data_fov_down',-25.0),semantic=sem)
        poses=synth.trajectory(4)
        try:
            f=R.Full(p); osl=O.Slam(p); res='ok'
            for t in range(4):
                sc=scene.scan(t,poses[t]); f.process_scan(*sc); osl.process_scan(*sc)
                if not np.array_equal(f.pose(),osl.pose()): res='t=%d pose differs %.2e'%(t,np.abs(f.pose()-osl.pose()).max()); break
                if f.map_size()!=osl.map.size(): res='t=%d size %d vs %d'%(t,f.map_size(),osl.map.size()); break
                try: surfel_fields_equal(f.map_download(),osl.map.download())
                except AssertionError as e: res='t=%d %s'%(t,str(e)[:120]); break
        except Exception as e: res='EXC '+str(e)[:160]
        print('sem' if sem else 'geo',{k:v for k,v in kw.items()},res,flush=True)
