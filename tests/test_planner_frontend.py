"""The batched BiRRT* front-end (shim/sqmp_birrt.hpp).  The same C++ planner is compiled twice: over
the CUDA library (product) and over the CPU oracle (test / baseline binary).  Near-neighbour ids are bit-exact and
collision verdicts equal outside the 1e-5 m band, so with the same seed both must take the SAME decisions: the trace
hash covers every nearest id, extend verdict, near count, chosen parent, rewire and connection."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from squirrel_motion_planner_b200 import synthetic as syn

sys.path.insert(0, os.path.join(ROOT, "tools"))
import planner_workdir as pw  # noqa: E402

SHIM = os.path.join(ROOT, "squirrel_motion_planner_b200", "shim")
DESC = os.path.join(ROOT, "squirrel_motion_planner_b200", "data", "robotino_plan.sqmpdesc")


def _workdir(tmp_path, fixture_data, world, starts, goals):
    lt, fx, spheres = fixture_data
    d = str(tmp_path)
    pw.write_world(d, world)
    pw.write_oracle_model(d, lt, spheres)
    pw.write_queries(d, starts, goals)
    return d


ORACLE_DIR = os.path.join(ROOT, "oracle")


def _run(binary, d, backend, iterations, seed=1):
    if binary == "sqmp_plan_oracle":      # the CPU twin is test infrastructure: built by and living under oracle/
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "planner"], check=True)
        where = ORACLE_DIR
    else:
        subprocess.run(["make", "-s", "-C", SHIM], check=True)
        where = SHIM
    r = subprocess.run([os.path.join(where, binary), DESC, d, backend, str(iterations), str(seed)], capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = [l.split() for l in r.stdout.splitlines() if l and not l.startswith("#")]
    return rows


def test_config1_single_query_on_cpu(tmp_path, fixture_data):
    """BASELINE.json configs[0]: one BiRRT* query in the synthetic box world on the CPU (reference-checker port)."""
    world = syn.box_world(g=256, n_boxes=64, seed=1234)
    d = _workdir(tmp_path, fixture_data, world, *pw.config1_queries(1))
    a = _run("sqmp_plan_oracle", d, "oracle", 400)
    b = _run("sqmp_plan_oracle", d, "oracle", 400)
    assert a == b                                            # seeded: identical decisions, identical trace
    row = a[0]
    assert row[1] == "1" and float(row[2]) < 40.0            # solved; path cost well under the 10000 sentinel
    assert int(row[3]) == 400 and int(row[8]) == 21 * int(row[7])   # 21 states per checked edge


def test_invalid_goal_is_rejected_like_init_planner(tmp_path, fixture_data):
    world = syn.box_world(g=128, n_boxes=8, seed=2)
    start, goal = pw.config1_queries(1)
    start[0, :2] = 1.0
    goal[0, :3] = (1.5, 1.5, 0.0)
    goal[0, 3:] = (0.29, 2.26, 0.19, -1.17, 0.53)            # hand inside the base: self collision (BS:351-358)
    d = _workdir(tmp_path, fixture_data, world, start, goal)
    assert _run("sqmp_plan_oracle", d, "oracle", 10)[0][1] == "invalid_endpoint"


@pytest.mark.gpu
def test_gpu_and_cpu_backends_take_identical_decisions(tmp_path, fixture_data):
    world = syn.box_world(g=256, n_boxes=64, seed=1234)
    rng = np.random.default_rng(5)
    s0, g0 = pw.config1_queries(1)
    starts = np.repeat(s0, 3, axis=0)
    goals = np.repeat(g0, 3, axis=0)
    starts[1, :2] = (3.5, -4.0); goals[1, :2] = (-4.0, 3.0)
    starts[2, :3] = (0.0, -4.5, 1.0); goals[2, 3:] += rng.uniform(-0.1, 0.1, 5)
    d = _workdir(tmp_path, fixture_data, world, starts, goals)
    cpu = _run("sqmp_plan_oracle", d, "oracle", 600)
    gpu = _run("sqmp_plan", d, "gpu", 600)
    assert len(cpu) == len(gpu) == 3
    for c, g in zip(cpu, gpu):
        assert c == g, "decision traces differ:\\n%s\\n%s" % (c, g)
    assert sum(r[1] == "1" for r in gpu) >= 2


def test_lockstep_scheduler_host_logic_on_the_cpu(tmp_path, fixture_data):
    """The lock-step scheduler (shim/sqmp_multi.hpp: fibers, request gathering, answer scatter) instantiated over the CPU
    oracle must print the sequential oracle runner's lines — batch sizes that do and do not divide the number of queries,
    an invalid goal in the middle of a batch."""
    world = syn.box_world(g=128, n_boxes=16, seed=21)
    rng = np.random.default_rng(13)
    s0, g0 = pw.config1_queries(1)
    n = 7
    starts, goals = np.repeat(s0, n, axis=0), np.repeat(g0, n, axis=0)
    starts[:, :2] = rng.uniform(-2.0, -1.0, size=(n, 2))
    goals[:, :2] = rng.uniform(1.0, 2.0, size=(n, 2))
    goals[3, 3:] = (0.29, 2.26, 0.19, -1.17, 0.53)          # self collision: rejected like BS:351-358
    d = _workdir(tmp_path, fixture_data, world, starts, goals)
    seq = _run("sqmp_plan_oracle", d, "oracle", 120)
    assert len(seq) == n and seq[3][1] == "invalid_endpoint"
    for batch in (3, 7, 64):
        r = subprocess.run([os.path.join(ORACLE_DIR, "sqmp_plan_oracle"), DESC, d, "oracle", "120", "1", "0", "0", str(n), str(batch)],
                           capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout + r.stderr
        rows = [l.split() for l in r.stdout.splitlines() if l and not l.startswith("#")]
        assert rows == seq, "batch %d" % batch


@pytest.mark.gpu
def test_lockstep_multi_query_runner_repeats_the_sequential_decisions(tmp_path, fixture_data):
    """shim/sqmp_plan_multi (B queries per GPU as fibers over the same planner code, one library call per request kind
    and round) must print the sequential runner's lines: same decisions, same trace hashes — for batch sizes that do and
    do not divide the number of queries, and with an invalid goal in the middle of a batch."""
    world = syn.box_world(g=256, n_boxes=64, seed=1234)
    rng = np.random.default_rng(11)
    s0, g0 = pw.config1_queries(1)
    n = 11
    starts, goals = np.repeat(s0, n, axis=0), np.repeat(g0, n, axis=0)
    starts[:, :2] += rng.uniform(-0.4, 0.4, size=(n, 2))
    goals[:, :2] += rng.uniform(-0.4, 0.4, size=(n, 2))
    goals[4, 3:] = (0.29, 2.26, 0.19, -1.17, 0.53)          # self collision: init_planner rejects it (BS:351-358)
    d = _workdir(tmp_path, fixture_data, world, starts, goals)
    seq = _run("sqmp_plan", d, "gpu", 300)
    assert len(seq) == n and seq[4][1] == "invalid_endpoint"
    for batch in (4, 11, 128):
        subprocess.run(["make", "-s", "-C", SHIM], check=True)
        r = subprocess.run([os.path.join(SHIM, "sqmp_plan_multi"), DESC, d, "gpu", "300", "1", "0", "0", str(n), str(batch)],
                           capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout + r.stderr
        rows = [l.split() for l in r.stdout.splitlines() if l and not l.startswith("#")]
        assert rows == seq, "batch %d" % batch
    # several host threads, each with its own context on the same GPU and the voxel fields shared (sqmp_share_voxel_grid)
    r = subprocess.run([os.path.join(SHIM, "sqmp_plan_multi"), DESC, d, "gpu", "300", "1", "0", "0", str(n), "3", "3"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert [l.split() for l in r.stdout.splitlines() if l and not l.startswith("#")] == seq


@pytest.mark.gpu
def test_config5_scene_queries_reach_the_optimisation_phase(tmp_path, fixture_data):
    """BASELINE.json configs[4] in small: planning queries in W512 (512^3 voxels, 2048 boxes, seed 4321) whose start / goal
    pairs come from the bench's own generator.  GPU backend and CPU twin must print the same lines (decision trace, costs,
    counters), the lock-step runner must repeat them, and the batch must exercise the whole path: at least half of the
    queries reach the optimisation phase (near-vertex search, choose-parent, rewire, informed samples)."""
    from squirrel_motion_planner_b200 import api, robot_desc as rd
    world = syn.box_world(g=512, n_boxes=2048, seed=4321)
    ctx = api.default_context(0)
    ctx.set_voxel_grid(world.bits, world.dims, world.origin + np.array([0.0, 0.0, rd.OCTREE_Z_OFFSET]), world.res)

    def in_collision(q):
        return np.unpackbits(ctx.check_states(q, api.CHECK_MAP | api.CHECK_SELF).view(np.uint8), bitorder="little")[:len(q)].astype(bool)
    n = 12
    starts, goals = pw.config5_queries(in_collision, n, seed=5)
    ctx.close()
    d = _workdir(tmp_path, fixture_data, world, starts, goals)
    iters = 250
    gpu = _run("sqmp_plan", d, "gpu", iters)
    r = subprocess.run([os.path.join(SHIM, "sqmp_plan_multi"), DESC, d, "gpu", str(iters), "1", "0", "0", str(n), "8"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert [l.split() for l in r.stdout.splitlines() if l and not l.startswith("#")] == gpu
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "planner"], check=True)
    r = subprocess.run([os.path.join(ORACLE_DIR, "sqmp_plan_oracle"), DESC, d, "oracle", str(iters), "1", "0", "0", "4"],
                       capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout + r.stderr
    cpu = [l.split() for l in r.stdout.splitlines() if l and not l.startswith("#")]
    assert cpu == gpu[:4], "decision traces differ"
    assert all(row[1] != "invalid_endpoint" for row in gpu)
    optimising = [row for row in gpu if int(row[5]) > 0]                    # near_calls > 0 <=> a solution existed
    assert len(optimising) >= n // 2 and sum(int(row[9]) for row in gpu) > 0 and sum(int(row[16]) for row in gpu) > 0
    assert all(row[1] == "1" for row in optimising)


def test_cpp_front_end_equals_independent_python_restatement(tmp_path, fixture_data, orc):
    """the decisions of shim/sqmp_birrt.hpp pinned by a SECOND restatement of the reference's iteration
    (tests/birrt_restatement.py: Python, one edge at a time with the reference's early exits, its own tree structure, the
    numpy twin of the ellipse set-up).  Three queries with different seeds, each long enough to find a solution and to
    spend most of its iterations in the optimisation phase (informed samples, choose-parent, rewires, rewires ON the
    solution path): decision trace, node counts, solution costs (total / revolute / prismatic) and counters must agree."""
    import birrt_restatement as br
    import oracle
    from squirrel_motion_planner_b200 import robot_desc as rd
    world = syn.box_world(g=160, n_boxes=14, seed=99, dims=(160, 160, 48))
    origin = world.origin + np.array([0.0, 0.0, rd.OCTREE_Z_OFFSET])
    arm_a, arm_b = pw.half_folded_arm(), np.array([-0.2, 1.2, 0.4, 0.9, 0.3])
    orc.set_disabled_links(0)
    orc.set_grid(world.bits, world.dims, origin, world.res)
    cand = syn.sample_configs(200, seed=5, xy_half=2.8)      # collision-free end points: the first six valid candidates
    cand[0::2, 3:], cand[1::2, 3:] = arm_a, arm_b
    ok = cand[~np.unpackbits(orc.check_states(cand, oracle.FLAG_SELF | oracle.FLAG_MAP).view(np.uint8), bitorder="little")[:len(cand)].astype(bool)]
    starts, goals = ok[0:3], ok[3:6]
    d = _workdir(tmp_path, fixture_data, world, starts, goals)
    iters, seed = 160, 11
    rows = _run("sqmp_plan_oracle", d, "oracle", iters, seed)
    lo = np.array([world.origin[0] + 1.0, world.origin[1] + 1.0, -3.14159265358979323846, -1.2, -1.7, -1.8, -2.4, -2.9])
    hi = np.array([world.origin[0] + world.dims[0] * world.res - 1.0, world.origin[1] + world.dims[1] * world.res - 1.0,
                   3.14159265358979323846, 1.5, 2.6, 1.8, 2.4, 2.9])
    phases = 0
    for k, row in enumerate(rows):
        p = br.Planner(orc, oracle.FLAG_SELF | oracle.FLAG_MAP, lo, hi, seed + k)
        assert p.init(starts[k], goals[k]) and row[1] != "invalid_endpoint"
        for _ in range(iters):
            p.iterate()
        s = p.stats
        assert int(row[1]) == int(p.solution) and int(row[13]) == p.trace, "query %d: decision traces differ" % k
        assert int(row[3]) == s["iterations"] and int(row[4]) == s["nn_calls"] and int(row[5]) == s["near_calls"]
        assert int(row[9]) == s["rewires"] and int(row[10]) == s["parents_chosen"]
        assert int(row[11]) == len(p.start_tree.nodes) and int(row[12]) == len(p.goal_tree.nodes)
        assert float(row[2]) == p.best[0] and float(row[14]) == p.best[1] and float(row[15]) == p.best[2]   # bit for bit
        assert int(row[16]) == s["informed_samples"] and int(row[17]) == s["solution_rewires"]
        if p.solution and s["rewires"] > 0 and s["informed_samples"] > iters // 3:
            phases += 1
    assert phases >= 2                                       # the comparison covers the optimisation phase, not only the search
