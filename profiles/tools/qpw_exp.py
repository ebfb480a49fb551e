import os, subprocess, sys, json
child = r'''
import sys, os, json; sys.path.insert(0, '.')
import numpy as np, torch
from quadrupedal_foothold_planner_b200 import _abi, synth
from quadrupedal_foothold_planner_b200.planner import Fhp
import bench
wl = bench.make_workload("c3")
m, cycles, params = wl["map"], wl["cycles"], wl["params"]
ctx = Fhp(0)
layers = [torch.from_numpy(np.ascontiguousarray(m.layers[n].T)).cuda() for n in wl["names"]]
ctx.set_map_device(m, wl["names"], [t.data_ptr() for t in layers], keepalive=layers)
out = {}
for nq in (1024, 2048, 4096, 8192, 16384, 32768):
    q = wl["queries"][:nq]; R = 4 + 4 * cycles
    parr, npar = Fhp._params_array(params)
    dq = torch.from_numpy(q.view(np.uint8)).cuda()
    dn = torch.empty(nq * R * 40, dtype=torch.uint8, device='cuda'); ds = torch.empty(nq * 8, dtype=torch.uint8, device='cuda')
    st = torch.cuda.Stream()
    for _ in range(3): ctx.plan_device(parr, npar, dq, nq, cycles, dn, None, ds, stream=st.cuda_stream)
    st.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(20): ctx.plan_device(parr, npar, dq, nq, cycles, dn, None, ds, stream=st.cuda_stream)
    e1.record(st); st.synchronize()
    out[nq] = round(e0.elapsed_time(e1) / 20, 4)
print(json.dumps(out))
'''
for sh in ("auto", "0", "1", "2", "3"):
    env = dict(os.environ)
    if sh != "auto": env["FHP_QPW_SHIFT"] = sh
    r = subprocess.run([sys.executable, '-c', child], capture_output=True, text=True, env=env)
    print("qpw shift", sh, r.stdout.strip() or r.stderr[-300:], flush=True)
