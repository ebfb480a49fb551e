import sys; sys.path.insert(0,'.')
import numpy as np
from quadrupedal_foothold_planner_b200 import _abi, synth
from quadrupedal_foothold_planner_b200.planner import Fhp
m = synth.rough_map(600, 600, 0.02, seed=3, start=(123,277))
q = synth.uniform_queries(512, (-5.0, 2.0), (-5.0, 5.0), seed=4)
p = _abi.yaml_params()
f=Fhp(0); f.set_map(m)
got=f.plan(p,q,12,centroid=True)
print("ok", got[2]['n_rows'].sum())
