import sys; sys.path.insert(0, '.')
import numpy as np, torch
from quadrupedal_foothold_planner_b200 import _abi, synth
from quadrupedal_foothold_planner_b200.planner import Fhp
import bench
nq = int(sys.argv[1]); wl = bench.make_workload("c3")
m, cycles, params = wl["map"], wl["cycles"], wl["params"]
ctx = Fhp(0)
layers = [torch.from_numpy(np.ascontiguousarray(m.layers[n].T)).cuda() for n in wl["names"]]
ctx.set_map_device(m, wl["names"], [t.data_ptr() for t in layers], keepalive=layers)
q = wl["queries"][:nq]; R = 4 + 4 * cycles
parr, npar = Fhp._params_array(params)
dq = torch.from_numpy(q.view(np.uint8)).cuda()
dn = torch.empty(nq * R * 40, dtype=torch.uint8, device='cuda'); ds = torch.empty(nq * 8, dtype=torch.uint8, device='cuda')
st = torch.cuda.Stream()
for _ in range(8): ctx.plan_device(parr, npar, dq, nq, cycles, dn, None, ds, stream=st.cuda_stream)
st.synchronize()
