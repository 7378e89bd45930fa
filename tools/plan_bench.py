"""Development timing of the planner front-end: sqmp_plan (GPU backend) vs sqmp_plan_oracle (CPU) on the same queries."""
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import planner_workdir as pw
from squirrel_motion_planner_b200 import robot_desc as rd, synthetic as syn
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "squirrel_motion_planner_b200", "shim"); DESC = os.path.join(ROOT, "squirrel_motion_planner_b200", "data", "robotino_plan.sqmpdesc")
lt, fx = rd.load_fixture(); spheres = [(n, tuple(c), r) for n, c, r in fx["spheres"]]
d = tempfile.mkdtemp()
w = syn.box_world(g=256, n_boxes=64, seed=1234)
pw.write_world(d, w); pw.write_oracle_model(d, lt, spheres)
nq = int(os.environ.get("NQ", "4")); iters = os.environ.get("ITERS", "2000")
s, g = pw.config1_queries(nq); pw.write_queries(d, s, g)
for binary, backend in (("sqmp_plan", "gpu"), ("sqmp_plan_oracle", "oracle")):
    t = time.time()
    r = subprocess.run([os.path.join(SHIM if backend == "gpu" else os.path.join(ROOT, "oracle"), binary), DESC, d, backend, iters, "1"], capture_output=True, text=True)
    print(backend, "rc", r.returncode, "wall %.2f s" % (time.time() - t)); print(r.stdout[-600:], r.stderr[-300:])
