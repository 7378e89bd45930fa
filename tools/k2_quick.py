"""Development aid: time K2 (sqmp_fk_dev) over a working set larger than L2 and check it against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import numpy as np, torch
import oracle
from squirrel_motion_planner_b200 import api, robot_desc as rd, synthetic as syn
ctx = api.default_context(0)
st = torch.cuda.current_stream().cuda_stream
S = 1 << 20
ins = [torch.from_numpy(syn.state_batch(S, seed=42 + i)).cuda() for i in range(4)]
outs = [torch.empty((S, 7), dtype=torch.float64, device="cuda") for _ in range(4)]
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
lt, fx = rd.load_fixture()
orc = oracle.Oracle(lt, [(n, tuple(c), r) for n, c, r in fx["spheres"]])
for prec in (64, 32):
    ctx.set_option("fk_precision", prec)
    for i in range(8):
        ctx.fk_dev(ins[i % 4], S, outs[i % 4], st)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(40):
        ctx.fk_dev(ins[i % 4], S, outs[i % 4], st)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 40
    got = outs[3][:4096].cpu().numpy()
    ref = orc.fk(ins[3][:4096].cpu().numpy())
    sgn = np.sign((got[:, 3:] * ref[:, 3:]).sum(1))[:, None]
    print("K2 fp%d: %.2f us per 2^20 states, %.3f of the measured HBM peak; max |pos err| %.2e, max |quat err| %.2e" % (
        prec, ms * 1e3, S * 120 / (ms * 1e-3) / 1e9 / peak, np.abs(got[:, :3] - ref[:, :3]).max(), np.abs(got[:, 3:] - sgn * ref[:, 3:]).max()))
