import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
from quadrupedal_foothold_planner_b200 import _abi, synth
from quadrupedal_foothold_planner_b200.planner import Fhp
def run(m, q, cycles, tag):
    ctx = Fhp(0); ctx.set_map(m)
    parr, npar = Fhp._params_array(_abi.yaml_params())
    nq = len(q); R = 4 + 4 * cycles
    hq = torch.from_numpy(q.view(np.uint8).copy()).pin_memory()
    raw = torch.zeros(nq * R * 24 + 128, dtype=torch.uint8).pin_memory(); hs = raw[(-raw.data_ptr()) % 128:]
    hm = torch.zeros(nq * 16, dtype=torch.int32).pin_memory()
    hst = torch.zeros(nq * 8, dtype=torch.uint8).pin_memory()
    for _ in range(3): ctx.plan_stream_raw(parr, npar, hq, nq, cycles, hs, hm, hst)
    t0 = time.perf_counter()
    for _ in range(10): ctx.plan_stream_raw(parr, npar, hq, nq, cycles, hs, hm, hst)
    dt = (time.perf_counter() - t0) / 10
    st = hst.numpy().view(_abi.STATUS_DTYPE)
    rows = int(st["n_rows"].sum())
    print(tag, "ms", round(dt * 1e3, 3), "rows", rows, "GB/s", round(rows * 20 / dt / 1e9, 1), "footholds/s %.3g" % (nq * cycles * 4 / dt))
    ctx.close()
m, q, c = synth.config_c3()
run(m, q, c, "rough")
f = synth.flat_map(4000, 4000, 0.02)
run(f, q, c, "flat")
