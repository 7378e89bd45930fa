"""CPU: the oracle (oracle/sqmp_oracle.c) against golden vectors made by independent code (tests/golden/make_golden.py).

The reference ships no tests or fixtures for this path ("parity unpinned", oracle/sqmp_oracle.h); these goldens pin
the oracle to scipy / numpy restatements computed straight from the reference's URDF constants.
"""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, unpack


def _quat_close(a, b, tol):
    return np.minimum(np.abs(a - b).max(1), np.abs(a + b).max(1)).max() < tol


def test_fk_matches_scipy_urdf_chain(orc):
    g = np.load(os.path.join(GOLDEN, "fk_scipy.npz"))
    ee = orc.fk(g["q"])
    assert np.abs(ee[:, :3] - g["ee"][:, :3]).max() < 1e-12
    assert _quat_close(ee[:, 3:], g["ee"][:, 3:], 1e-12)
    assert np.abs(np.linalg.norm(ee[:, 3:], axis=1) - 1).max() < 1e-12


def test_link_frames_match_scipy(orc, fixture_data):
    lt = fixture_data[0]
    g = np.load(os.path.join(GOLDEN, "fk_scipy.npz"))
    for s in range(0, len(g["q"]), 16):
        fr = orc.update_transforms(g["q"][s])
        for n in g["frame_names"]:
            assert np.abs(fr[lt.index(str(n))] - g["frame_" + str(n)][s]).max() < 1e-12, n


def test_root_offset_and_frozen_joints(orc, fixture_data):
    lt = fixture_data[0]
    fr = orc.update_transforms(np.zeros(8))
    assert fr[0, 11] == 0.02 and np.all(fr[0, 9:11] == 0)          # CC:199-201
    assert abs(fr[lt.index("base_neck_link"), 11] - (0.676 + 0.02)) < 1e-15   # URDF base_neck_joint, frozen


def test_collision_clearance_matches_numpy_bruteforce(orc):
    g = np.load(os.path.join(GOLDEN, "collision_numpy.npz"))
    orc.set_disabled_links(0)
    orc.set_grid(g["bits"], g["dims"], g["origin"], float(g["res"]))
    mask, cm, _ = orc.check_states(g["q"], oracle.FLAG_MAP, clearance=True, cap=10.0)
    assert np.abs(cm - g["clearance"]).max() < 1e-12
    hit = unpack(mask, len(g["q"]))
    assert np.array_equal(hit, g["clearance"] <= 0.0)
    assert 0.1 < hit.mean() < 0.9
    # the early-exit path gives the same verdicts
    assert np.array_equal(unpack(orc.check_states(g["q"], oracle.FLAG_MAP), len(g["q"])), hit)
    # getCollisions (CC:610-631): per-link map hits
    for s in range(0, len(g["q"]), 8):
        lm, _ = orc.get_collisions(g["q"][s])
        assert np.array_equal(lm, g["per_link"][s])


def test_disabled_links_and_flags(orc):
    g = np.load(os.path.join(GOLDEN, "collision_numpy.npz"))
    orc.set_grid(g["bits"], g["dims"], g["origin"], float(g["res"]))
    q = g["q"]
    assert not orc.is_in_collision(q[0], False, False)               # CC:109-110
    orc.set_disabled_links((1 << 64) - 1)
    assert not unpack(orc.check_states(q, oracle.FLAG_MAP), len(q)).any()
    orc.set_disabled_links(0)


def test_nearest_and_radius_match_numpy(orc):
    g = np.load(os.path.join(GOLDEN, "nn_numpy.npz"))
    ids = oracle.nearest(g["nodes"], g["queries"])
    assert np.array_equal(ids, g["nearest"])
    assert ids[0] == 50                                               # duplicates at 50, 100, 2000: lowest id wins
    off = 0
    for k, x in enumerate(g["queries"]):
        n = int(g["near_count"][k])
        want = g["near_flat"][off:off + n]
        off += n
        got = oracle.near_vertices(g["nodes"], g["cost"], x, -1, 1.5)
        assert np.array_equal(got, want)
        if n:
            ex = int(want[0])
            got_ex = oracle.near_vertices(g["nodes"], g["cost"], x, ex, 1.5)
            assert np.array_equal(got_ex, want[1:])                   # BS:4295 excludes x_new itself
    low, high, cnt = oracle.near_topk(g["nodes"], g["cost"], g["queries"], None, 1.5, 20)
    assert np.array_equal(cnt, g["near_count"])


def test_nearest_quirks():
    far = np.full((5, 8), 1e6)
    assert oracle.nearest(far, np.zeros((1, 8)))[0] == 0              # nothing below 10000.0: id stays 0 (BS:4084-4090)
    nodes = np.zeros((3, 8)); nodes[0, 0] = 20000.0; nodes[1, 0] = 3.0; nodes[2, 0] = 3.0
    assert oracle.nearest(nodes, np.zeros((1, 8)))[0] == 1


def test_interpolation_matches_numpy(orc):
    g = np.load(os.path.join(GOLDEN, "interp_numpy.npz"))
    for e in range(len(g["start"])):
        got = oracle.interpolate(g["start"][e], g["end"][e], 20)
        assert np.array_equal(got, g["states"][e * 21:(e + 1) * 21])   # bit-exact
        assert np.array_equal(got[0], g["start"][e])


def test_is_edge_valid_last_valid_idx(orc):
    g = np.load(os.path.join(GOLDEN, "collision_numpy.npz"))
    orc.set_grid(g["bits"], g["dims"], g["origin"], float(g["res"]))
    from squirrel_motion_planner_b200 import synthetic as syn
    a, b = syn.edge_batch(200, seed=9, xy_half=1.0)
    mask, first = orc.check_edges(a, b, 20, oracle.FLAG_MAP)
    bits = unpack(mask, 200 * 21).reshape(200, 21)
    seen = set()
    for e in range(200):
        traj = oracle.interpolate(a[e], b[e], 20)
        ok, last = orc.is_edge_valid(traj, False, True)
        bad = np.nonzero(bits[e])[0]
        assert ok == (bad.size == 0)
        assert first[e] == (bad[0] if bad.size else -1)
        assert last == (20 if ok else max(bad[0] - 1, 0))             # BS:6878-6895
        seen.add(int(first[e]))
    assert -1 in seen and 0 in seen and len(seen) > 3


def test_edge_cost_and_step(orc):
    from squirrel_motion_planner_b200 import synthetic as syn
    a = syn.sample_configs(50, seed=1)
    r = syn.sample_configs(50, seed=2)
    r[:5, :2] = a[:5, :2] + 1e-4          # prismatic target already reached
    r[5:10, 2:] = a[5:10, 2:]             # revolute target already reached
    r[10:12] = a[10:12]
    stepped = syn.step_towards(a, r, 0.6)
    for i in range(50):
        reached, x = oracle.step_towards(a[i], r[i], syn.IS_REVOLUTE, 0.6)
        assert np.array_equal(x, stepped[i])                          # the vectorised host version is bit-exact
        assert reached == (np.array_equal(x, r[i]))
    traj = oracle.interpolate(a[20], stepped[20], 20)
    c = oracle.edge_cost(traj, syn.IS_REVOLUTE)
    seg = np.diff(traj, axis=0)
    assert abs(c[0] - np.sqrt((seg ** 2).sum(1)).sum()) < 1e-12
    assert abs(c[1] - np.sqrt((seg[:, 2:] ** 2).sum(1)).sum()) < 1e-12
    assert abs(c[2] - np.sqrt((seg[:, :2] ** 2).sum(1)).sum()) < 1e-12


def test_published_poses_against_the_conservative_model(orc, fixture_data):
    """Sanity of the sphere model against poses the real robot executes (config/parameters.yaml:37 folded arm;
    folding_poses_tuw-robotino2.yaml key-frames 2..24).  The 20 key-frames with the arm away from the platform are self-collision
    free.  At the folded end of the trajectory (the first four) the hand rests on the shell: there the model reports contact
    with the front shell box / the platform drum only — and it must, because the reference's own URDF box (0.55 x 0.55 x 0.17,
    robotino_plan.urdf:323-330) already contains the finger joints at those poses (checked below from the link frames alone)."""
    lt, fx = fixture_data[0], fixture_data[1]
    poses = [fx["pose_folded_arm"]] + fx["folding_poses"][1:]
    q = np.array([[0.3, -0.2, 0.7] + list(a) for a in poses])
    hit = unpack(orc.check_states(q, oracle.FLAG_SELF), len(q))
    assert not hit[4:].any()
    base = {"shell_base_link_front", "base_body_link"}
    for i in np.nonzero(hit[:4])[0]:
        _, ph = orc.get_collisions(q[i])
        pairs = [(lt.names[a], lt.names[b]) for a, b in zip(*np.nonzero(ph)) if a < b]
        assert pairs and all(a in base or b in base for a, b in pairs), pairs
    fr = orc.update_transforms(q[1])                       # key-frame 2
    box, tip = fr[lt.index("shell_base_link_front")], fr[lt.index("hand_middle_finger_upper_link")]
    local = box[:9].reshape(3, 3).T @ (tip[9:] - box[9:])    # finger joint in the box's collision frame
    assert np.all(np.abs(local) < np.array([0.275, 0.275, 0.085]))


def test_sampler_is_the_documented_counter_stream(orc):
    """orc_sample_configs (sampleJointConfig_JntArray BS:3832-3878 / getRandomConf CL:1120-1188) against an independent
    pure-Python restatement of the splitmix64 counter stream, the uniform map and the 0.0 <= ee_z acceptance rule."""
    lo = np.array([-5.0, -4.0, -np.pi, -1.2, -1.7, -1.8, -2.4, -2.9])
    hi = np.array([5.0, 6.0, np.pi, 1.5, 2.6, 1.8, 2.4, 2.9])
    M, G = (1 << 64) - 1, 0x9E3779B97F4A7C15

    def out(seed, k):
        z = (seed + (k + 1) * G) & M
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return z ^ (z >> 31)

    seed, first, n, T = 0xDEADBEEFCAFE, 1000, 300, 5
    q, tries = orc.sample_configs(seed, first, n, lo, hi, T)
    for i in range(n):
        used = -T
        for t in range(T):
            k0 = ((first + i) * T + t) * 8
            cand = np.array([lo[j] + (hi[j] - lo[j]) * ((out(seed, k0 + j) >> 11) * (1.0 / 9007199254740992.0)) for j in range(8)])
            if orc.fk(cand[None])[0, 2] >= 0.0:
                used = t + 1
                break
        assert tries[i] == used and np.array_equal(q[i], cand)
    assert np.all(q >= lo) and np.all(q <= hi)
    # order-free: a batch is the concatenation of its parts; a different seed is a different stream
    qa, _ = orc.sample_configs(seed, first, 100, lo, hi, T)
    qb, _ = orc.sample_configs(seed, first + 100, 200, lo, hi, T)
    assert np.array_equal(q, np.concatenate([qa, qb]))
    assert not np.array_equal(orc.sample_configs(seed + 1, first, 10, lo, hi, T)[0], q[:10])
    # a box whose arm poses dive below the platform: rejections happen and are reported
    lo2, hi2 = lo.copy(), hi.copy()
    lo2[3:] = [1.3, -1.7, -1.8, -0.4, -2.9]; hi2[3:] = [1.5, -1.0, -0.8, 0.8, 2.9]
    q2, t2 = orc.sample_configs(3, 0, 200, lo2, hi2, 4)
    ee = orc.fk(q2)[:, 2]
    assert np.all((t2 > 0) == (ee >= 0.0))
    assert (t2 > 1).any() and (t2 < 0).any()                 # some samples needed a redraw, some ran out of tries


def test_informed_sampler_is_the_documented_rule(orc):
    """orc_sample_informed (sampleJointConfigfromEllipse_JntArray BS:3607-3829 over the ellipse of BS:3472-3604) against a
    pure-Python restatement of the counter stream and of the map ball -> ellipse, and against the geometry it must have:
    every sample lies inside the ellipse of its joint class, above the platform, inside the environment."""
    from squirrel_motion_planner_b200 import informed
    is_rev = [0, 0, 1, 1, 1, 1, 1, 1]
    start = np.array([-2.0, -1.5, 0.3, -0.17, 1.96, 0.19, 0.65, 0.53])
    goal = np.array([1.5, 1.0, 1.2, -0.2, 1.2, 0.4, 0.9, 0.3])
    el = informed.ellipse_init(start, goal, is_rev)
    for cls in ("rev", "pri"):                               # the rotation C: orthogonal, det +1, first column (goal - start) / c_min
        Cm, e = el[cls]["C"], el[cls]
        assert np.abs(Cm @ Cm.T - np.eye(len(Cm))).max() < 1e-15 and abs(np.linalg.det(Cm) - 1.0) < 1e-14
        assert np.abs(Cm[:, 0] * e["c_min"] - (goal - start)[e["idx"]]).max() < 1e-15
        assert np.array_equal(e["center"], ((start + goal) / 2.0)[e["idx"]])
    th = np.arctan2(goal[1] - start[1], goal[0] - start[0])   # two joints: THE rotation by the heading start -> goal
    assert np.abs(el["pri"]["C"] - np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])).max() < 1e-15
    c_rev, c_pri = 1.6 * el["rev"]["c_min"], 1.25 * el["pri"]["c_min"]
    lo = np.array([-5.0, -4.0, -np.pi, -1.2, -1.7, -1.8, -2.4, -2.9])
    hi = np.array([5.0, 6.0, np.pi, 1.5, 2.6, 1.8, 2.4, 2.9])
    es = informed.ctypes_struct(el, c_rev, c_pri, (lo[0], hi[0]), (lo[1], hi[1]))
    M, G = (1 << 64) - 1, 0x9E3779B97F4A7C15

    def out(seed, k):
        z = (seed + (k + 1) * G) & M
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return z ^ (z >> 31)

    seed, first, n, T = 0xABCDEF987654, 77, 200, 6
    q, tries = orc.sample_informed(seed, first, n, lo, hi, es, T)
    for i in range(n):
        used = -T
        for t in range(T):
            k0 = ((first + i) * T + t) * 9
            ball = np.array([lo[j] + (hi[j] - lo[j]) * ((out(seed, k0 + j) >> 11) * (1.0 / 9007199254740992.0)) for j in range(8)])
            scale = (out(seed, k0 + 8) >> 11) * (1.0 / 9007199254740992.0)
            cand = []
            for cls, cb in (("pri", c_pri), ("rev", c_rev)):
                e = el[cls]
                v = ball[e["idx"]]
                s2 = 0.0
                for x in v:
                    s2 += float(x) * float(x)
                l0, l1 = informed.ellipse_L(cb, e["c_min"])
                y = [(l0 if a == 0 else l1) * (scale * (float(v[a]) / np.sqrt(s2))) for a in range(len(v))]
                for r in range(len(v)):
                    acc = 0.0
                    for a in range(len(v)):
                        acc += float(e["C"][r, a]) * y[a]
                    cand.append(acc + float(e["center"][r]))
            cand = np.array(cand)
            if orc.fk(cand[None])[0, 2] >= 0.0 and lo[0] < cand[0] < hi[0] and lo[1] < cand[1] < hi[1]:
                used = t + 1
                break
        assert tries[i] == used and np.array_equal(q[i], cand)
    ok = tries > 0
    assert ok.mean() > 0.9
    for cls, cb in (("rev", c_rev), ("pri", c_pri)):         # inside the ellipse: |L^-1 C^T (x - centre)| <= 1
        e = el[cls]
        l0, l1 = informed.ellipse_L(cb, e["c_min"])
        u = (q[ok][:, e["idx"]] - e["center"]) @ e["C"] / np.array([l0] + [l1] * (len(e["idx"]) - 1))
        assert np.linalg.norm(u, axis=1).max() <= 1.0 + 1e-12
    assert (orc.fk(q[ok])[:, 2] >= 0.0).all()
    qa, _ = orc.sample_informed(seed, first, 50, lo, hi, es, T)   # order-free like the uniform stream
    qb, _ = orc.sample_informed(seed, first + 50, 150, lo, hi, es, T)
    assert np.array_equal(q, np.concatenate([qa, qb]))


def test_edge_cost_three_way_with_weights():
    """compute_edge_cost_interpolation (BS:4162-4242): total uses m_edge_cost_weights, the other two do not."""
    from squirrel_motion_planner_b200 import synthetic as syn
    a, b = syn.edge_batch(20, seed=3)
    w = np.array([2.0, 0.5, 1.0, 3.0, 1.0, 0.25, 1.0, 4.0])
    for i in range(20):
        traj = oracle.interpolate(a[i], b[i], 20)
        c = oracle.edge_cost(traj, syn.IS_REVOLUTE, w)
        seg = np.diff(traj, axis=0)
        assert abs(c[0] - np.sqrt((seg ** 2 * w).sum(1)).sum()) < 1e-12
        assert abs(c[1] - np.sqrt((seg[:, 2:] ** 2).sum(1)).sum()) < 1e-12 and abs(c[2] - np.sqrt((seg[:, :2] ** 2).sum(1)).sum()) < 1e-12
