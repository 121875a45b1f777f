import sys, os, faulthandler
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
faulthandler.enable(); faulthandler.dump_traceback_later(40, exit=True)
import numpy as np
from semantic_suma_b200 import api
from helpers import scans, sized
pp = api.default_params(**sized(900), max_iterations=8, stopping_threshold=0.0, delta=0.0)
sc, _ = scans(900, n=4)
sl = api.SurfelMapping(pp)
for i, s in enumerate(sc):
    sl.processScan(*s); print("scan", i, sl.getCurrentPose()[0, 3], flush=True)
print("ok")
