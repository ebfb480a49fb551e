"""usage: python scratch/kbench.py lib1.so lib2.so ... [--wl c3,c5]  — resident plan time per library (subprocess each)"""
import os, subprocess, sys, json
libs = [a for a in sys.argv[1:] if not a.startswith('--')]
wls = 'c3'
for a in sys.argv[1:]:
    if a.startswith('--wl='): wls = a[5:]
child = r'''
import sys, os, time, json; sys.path.insert(0, '.')
import numpy as np, torch
from quadrupedal_foothold_planner_b200 import _abi, synth
from quadrupedal_foothold_planner_b200.planner import Fhp
import bench
out = {}
for name in sys.argv[1].split(','):
    wl = bench.make_workload(name)
    m, q, cycles, params = wl["map"], wl["queries"], wl["cycles"], wl["params"]
    ctx = Fhp(0)
    layers = [torch.from_numpy(np.ascontiguousarray(m.layers[n].T)).cuda() for n in wl["names"]]
    ctx.set_map_device(m, wl["names"], [t.data_ptr() for t in layers], keepalive=layers)
    nq = len(q); R = 4 + 4 * cycles
    parr, npar = Fhp._params_array(params)
    dq = torch.from_numpy(q.view(np.uint8)).cuda()
    dn = torch.empty(nq * R * 40, dtype=torch.uint8, device='cuda'); ds = torch.empty(nq * 8, dtype=torch.uint8, device='cuda')
    st = torch.cuda.Stream()
    best = []
    for rep in range(3):
        for _ in range(3): ctx.plan_device(parr, npar, dq, nq, cycles, dn, None, ds, stream=st.cuda_stream)
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(20): ctx.plan_device(parr, npar, dq, nq, cycles, dn, None, ds, stream=st.cuda_stream)
        e1.record(st); st.synchronize()
        best.append(e0.elapsed_time(e1) / 20)
    chk = int(ds.cpu().numpy().view(_abi.STATUS_DTYPE)["n_rows"].astype(np.int64).sum())
    out[name] = (round(min(best), 4), round(sorted(best)[1], 4), chk)
    ctx.close(); del layers, dq, dn, ds
print(json.dumps(out))
'''
for lib in libs:
    env = dict(os.environ, FHP_LIB=os.path.abspath(lib))
    r = subprocess.run([sys.executable, '-c', child, wls], capture_output=True, text=True, env=env)
    print(os.path.basename(lib), r.stdout.strip() or r.stderr[-400:], flush=True)
