import sys, os, ctypes as C
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from semantic_suma_b200 import api, synth
w=2048
pp = api.default_params(data_width=w, model_width=w, max_iterations=10, stopping_threshold=0.0, delta=0.0)
ctx = api.Context(pp)
sc = synth.Scene(width=w, height=64); poses = synth.trajectory(2)
f0, f1 = api.Frame(ctx, w, 64), api.Frame(ctx, w, 64)
pre = api.Preprocessing(ctx)
pre.process(sc.scan(0, poses[0])[0], f0); pre.process(sc.scan(1, poses[1])[0], f1)
obj = api.Frame2Model(ctx); obj.setData(f1, f0)
gn = api.LieGaussNewton(ctx)
for rep in range(3): gn.minimize(obj, np.eye(4))
L = api.lib(); L.sb_debug_icp_trace.argtypes=[C.c_void_p, C.c_void_p]
t = np.zeros(256, np.uint64)
print('rc', L.sb_debug_icp_trace(ctx.h, C.c_void_p(t.ctypes.data)))
t = t.reshape(16,16).astype(np.int64)
base = t[0,0]
names = {0:'start',1:'accum',2:'reduce',3:'ticket',8:'L:begin',9:'L:summed',12:'L:ldlt',13:'L:solve',10:'L:gn',11:'L:published'}
for it in range(11):
    row = t[it]
    if row[0]==0: break
    print(it, ' '.join('%s=%.2f'%(names[k], (row[k]-base)/1000.) for k in [0,1,2,3,8,9,12,13,10,11] if row[k]))
import torch, time
torch.cuda.synchronize(); 
s=torch.cuda.ExternalStream(ctx.stream()); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record(s)
for rep in range(20): gn.minimize(obj, np.eye(4))
e1.record(s); torch.cuda.synchronize(); print('avg minimize ms', e0.elapsed_time(e1)/20, 'blocks env', os.environ.get('SUMA_B200_ICP_BLOCKS'))
