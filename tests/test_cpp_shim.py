"""The C++ host shim (squirrel_motion_planner_b200/shim/sqmp_shim.hpp) — the reference-named classes a maintainer
drops into birrt_star_algorithm — driven by shim/test_shim.cpp the way birrt_star.cpp drives the reference classes,
and compared with the CPU oracle on the same inputs.  CPU part: it builds, and without a GPU it fails closed."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from conftest import ROOT
from squirrel_motion_planner_b200 import robot_desc as rd, synthetic as syn

SHIM_DIR = os.path.join(ROOT, "squirrel_motion_planner_b200", "shim")
BIN = os.path.join(SHIM_DIR, "test_shim")
DESC = os.path.join(ROOT, "squirrel_motion_planner_b200", "data", "robotino_plan.sqmpdesc")
BAND = 1e-5


def _build():
    subprocess.run(["make", "-s", "-C", SHIM_DIR], check=True)
    assert os.path.exists(BIN)


def test_desc_file_matches_the_compiled_robot(fixture_data):
    """data/robotino_plan.sqmpdesc (what the C++ side loads) is the export of the same compile_robot() the tests use."""
    lt, fx, spheres = fixture_data
    cr = rd.compile_robot(lt, spheres)
    toks = open(DESC).read().split("\n")
    assert toks[0] == "sqmpdesc 2" and toks[1] == "n_seg %d" % cr.n_seg
    sph = [l.split() for l in toks if l.startswith("sphere ")]
    assert len(sph) == cr.n_spheres
    assert np.array_equal(np.array([[float(v) for v in l[3:7]] for l in sph]),
                          np.column_stack([cr.sph_center, cr.sph_radius]))
    rows = {int(l.split()[1]): int(l.split()[2], 16) for l in toks if l.startswith("link_pairs ")}
    assert rows == {i: int(v) for i, v in enumerate(cr.link_pair_mask) if int(v)}


def test_shim_builds_and_fails_closed_without_gpu():
    _build()
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present: the closed-failure path needs a box without one")
    r = subprocess.run([BIN, DESC, "/nonexistent"], capture_output=True, text=True)
    assert r.returncode == 3 and "isInCollision -> 1" in r.stdout          # CC:106-107


@pytest.mark.gpu
def test_cpp_shim_against_oracle(tmp_path, fixture_data, orc, small_world):
    _build()
    lt, fx, spheres = fixture_data
    w, origin = small_world                      # origin already carries the -0.02 octree offset (CC:87)
    d = str(tmp_path)
    q = syn.state_batch(21 * 60, seed=91, xy_half=2.2)
    nodes, cost = syn.random_tree(1500, seed=92, cost_seed=93, xy_half=1.5)
    cost[::9] = cost[4]
    queries = syn.sample_configs(24, seed=94, xy_half=1.5)
    q.tofile(os.path.join(d, "states.f64"))
    nodes.tofile(os.path.join(d, "tree_nodes.f64"))
    cost.tofile(os.path.join(d, "tree_cost.f64"))
    queries.tofile(os.path.join(d, "queries.f64"))
    w.bits.tofile(os.path.join(d, "grid_bits.u32"))
    with open(os.path.join(d, "grid_meta.txt"), "w") as f:     # the shim applies the -0.02 itself: hand it the raw origin
        f.write("%d %d %d %.17g %.17g %.17g %.17g\n" % (*w.dims, *w.origin, w.res))
    disabled = ["head_link", "arm_link1", "hand_wrist_link"]
    with open(os.path.join(d, "disabled_links.txt"), "w") as f:
        f.write("\n".join(disabled) + "\n")
    r = subprocess.run([BIN, DESC, d], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr

    def rd_u8(name):
        return np.fromfile(os.path.join(d, name), dtype=np.uint8).astype(bool)

    S = len(q)
    orc.set_disabled_links(0)
    orc.set_grid(w.bits, w.dims, origin, w.res)
    assert not rd_u8("out_no_map.u8").any()                                   # CC:556-557: no map yet -> free

    def check(got, flags):
        ref_mask, cm, cs = orc.check_states(q, flags, clearance=True, nthreads=8)
        ref = np.unpackbits(ref_mask.view(np.uint8), bitorder="little")[:S].astype(bool)
        bad = (got != ref) & (np.abs(np.minimum(cm, cs)) > BAND)
        assert not bad.any()
        return ref
    check(rd_u8("out_verdict.u8"), oracle.FLAG_MAP | oracle.FLAG_SELF)
    ref_map = check(rd_u8("out_verdict_map.u8"), oracle.FLAG_MAP)
    assert 0.05 < ref_map.mean() < 0.95

    edges = np.fromfile(os.path.join(d, "out_edges.i32"), dtype=np.int32).reshape(-1, 3)
    for e, (ok1, ok2, last) in enumerate(edges):
        want_ok, want_last = orc.is_edge_valid(q[e * 21:(e + 1) * 21], False, True)
        assert bool(ok1) == bool(ok2) == want_ok and last == want_last          # BS:6864-6895

    orc.set_disabled_links(rd.link_mask(lt, disabled))
    check(rd_u8("out_verdict_disabled.u8"), oracle.FLAG_MAP)
    orc.set_disabled_links(0)

    lines = open(os.path.join(d, "out_collisions.txt")).read().split("\n")
    s0 = int(lines[0].split()[1])
    lm, ph = orc.get_collisions(q[s0])
    assert sorted(l.split()[1] for l in lines if l.startswith("map ")) == sorted(lt.names[i] for i in np.nonzero(lm)[0])
    want_pairs = sorted((lt.names[a], lt.names[b]) for a, b in zip(*np.nonzero(np.triu(ph | ph.T))))
    assert sorted(tuple(l.split()[1:3]) for l in lines if l.startswith("self ")) == want_pairs

    nn = np.fromfile(os.path.join(d, "out_nearest.i32"), dtype=np.int32)
    assert np.array_equal(nn, oracle.nearest(nodes, queries))                  # BS:4076-4133, bit-exact
    near = np.fromfile(os.path.join(d, "out_near.i32"), dtype=np.int32)
    off = 0
    for k in range(len(queries)):
        n = int(near[off]); got = near[off + 1:off + 1 + n]; off += 1 + n
        want = oracle.near_vertices(nodes, cost, nodes[nn[k]], int(nn[k]), 1.5)
        assert n == len(want)
        assert np.array_equal(got[:20], want[:20]) and np.array_equal(got[-20:], want[-20:])   # the two walks of the callers
        if n <= 40:
            assert np.array_equal(got, want)
    moved = nodes.copy(); moved[3] = queries[0]
    after = np.fromfile(os.path.join(d, "out_nearest_after_update.i32"), dtype=np.int32)
    assert after[0] == oracle.nearest(moved, queries[:1])[0] == 3              # BS:5173-5193

    fk = np.fromfile(os.path.join(d, "out_fk.f64")).reshape(-1, 7)
    ref = orc.fk(q[:len(fk)])
    assert np.abs(fk[:, :3] - ref[:, :3]).max() < 1e-5
    assert np.minimum(np.abs(fk[:, 3:] - ref[:, 3:]).max(1), np.abs(fk[:, 3:] + ref[:, 3:]).max(1)).max() < 1e-5

    samples = np.fromfile(os.path.join(d, "out_samples.f64")).reshape(-1, 8)      # sampleJointConfig_Vector, BS:3881-3930
    lo = np.array([-5.0, -4.0, -3.0, -1.2, -1.7, -1.8, -2.4, -2.9]); hi = np.array([5.0, 6.0, 3.0, 1.5, 2.6, 1.8, 2.4, 2.9])
    want, _ = orc.sample_configs(20240607, 0, 600, lo, hi, 64)
    assert np.array_equal(samples, want)

    # ---- the reference's exact signatures over a lazily mirrored host tree (BH:322-325; no slot, no mirror calls) ----
    lit = np.fromfile(os.path.join(d, "out_literal.i32"), dtype=np.int32)
    half = len(cost) // 2
    t_nodes, t_cost = nodes.copy(), cost.copy()
    t_nodes[3] = queries[0]                                                     # (node 3 was moved by the update test above)
    phases = [(t_nodes[:half].copy(), t_cost[:half].copy()), (t_nodes.copy(), t_cost.copy())]
    for n in range(5, len(cost), 7):                                            # the rewires + cost changes of test_shim.cpp
        t_nodes[n] = queries[n % 16]
        t_cost[n] *= 0.5
    phases += [(t_nodes.copy(), t_cost.copy()), (t_nodes[:40].copy(), t_cost[:40].copy())]
    off = 0
    for pn, pc in phases:
        for k in range(24):
            want_nn = int(oracle.nearest(pn, queries[k:k + 1])[0])
            assert lit[off] == want_nn
            n = int(lit[off + 1]); got = lit[off + 2:off + 2 + n]; off += 2 + n
            want = oracle.near_vertices(pn, pc, pn[want_nn], want_nn, 1.5)
            assert n == len(want) and np.array_equal(got[:20], want[:20]) and np.array_equal(got[-20:], want[-20:])
    assert off == len(lit)
    # isConfigValid(JntArray-like) (BS:6901-6908), compute_edge_cost_interpolation with weights (BS:4162-4242)
    jv = rd_u8("out_jntarray_valid.u8")
    first = q[0:21 * 40:21]
    ref_mask, cm, cs = orc.check_states(first, oracle.FLAG_MAP | oracle.FLAG_SELF, clearance=True)
    ref = np.unpackbits(ref_mask.view(np.uint8), bitorder="little")[:len(first)].astype(bool)
    assert not np.any((jv[:len(first)] == ref) & (np.abs(np.minimum(cm, cs)) > BAND))      # valid == not in collision
    ec = np.fromfile(os.path.join(d, "out_edge_cost.f64")).reshape(-1, 3)
    wts = np.ones(8); wts[1] = 0.5; wts[4] = 3.0
    for e in range(len(ec)):
        assert np.array_equal(ec[e], oracle.edge_cost(q[e * 21:(e + 1) * 21], syn.IS_REVOLUTE, wts))       # bit for bit
    # informed sampler through jointConfigEllipseInitialization / sampleJointConfigfromEllipse_Vector
    from squirrel_motion_planner_b200 import informed
    el = informed.ellipse_init(q[0], q[500], syn.IS_REVOLUTE)
    lo[2], hi[2] = -3.0, 3.0
    es = informed.ctypes_struct(el, 9.0, 14.0, (lo[0], hi[0]), (lo[1], hi[1]))
    want_inf, _ = orc.sample_informed((424242 + 0x5851f42d4c957f2d) & ((1 << 64) - 1), 0, 40, lo, hi, es, 64)
    got_inf = np.fromfile(os.path.join(d, "out_informed.f64")).reshape(-1, 8)
    assert np.array_equal(got_inf, want_inf)

