"""CPU: host-side logic — robot description folding, synthetic inputs, C-ABI exports, sharding over gloo."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, unpack
from squirrel_motion_planner_b200 import api, robot_desc as rd, sharding, sphere_model, synthetic as syn


def test_compile_robot_structure(fixture_data):
    lt, fx, spheres = fixture_data
    rigid = [tuple(p) for p in lt.extras["rigid_pairs_disabled"]]
    assert lt.n == 39 and int(lt.shaped.sum()) == 27 and len(lt.disabled_pairs) - len(rigid) == 245   # URDF / SRDF:26-271
    grp = rd.rigid_group_of_links(lt)   # pairs of links that are rigid to each other are out of the matrix (robot_desc.disable_rigid_pairs)
    assert rigid and all(grp[lt.index(a)] == grp[lt.index(b)] for a, b in rigid)
    assert ("hokuyo_link", "shell_base_link_front") in rigid          # the URDF cylinder lies inside the URDF box
    cr = rd.compile_robot(lt, spheres)
    assert cr.n_seg == 8 and list(cr.seg_qidx) == list(range(8))
    assert list(cr.seg_type) == [2, 2, 1, 1, 1, 1, 1, 1]
    assert np.array_equal(cr.seg_axis[3], [0, -1, 0]) and np.array_equal(cr.seg_axis[4], [0, 1, 0])   # URDF:615-650
    assert cr.root_offset[2] == 0.02 and cr.self_const_collision == 0
    assert cr.n_spheres == len(spheres) <= api.MAX_SPHERES
    assert np.all(np.diff(cr.sph_seg) >= 0)
    assert set(lt.names[i] for i in cr.sph_link) == set(n for n, s in zip(lt.names, lt.shaped) if s)
    m = lt.self_pair_matrix()
    for a in range(lt.n):
        assert int(cr.link_pair_mask[a]) == sum(1 << b for b in range(lt.n) if m[a, b])


def test_compiled_chain_equals_link_walk(fixture_data, orc):
    """The 8-segment folding reproduces the 39-link walk (CC:519-539) for every sphere centre."""
    lt, fx, spheres = fixture_data
    cr = rd.compile_robot(lt, spheres)
    q = syn.sample_configs(20, seed=3)
    for s in range(len(q)):
        fr = orc.update_transforms(q[s])
        T = rd.frame(p=cr.root_offset)
        seg_T = {-1: T.copy()}
        for k in range(cr.n_seg):
            T = T @ rd.frame(cr.seg_pre[k][:9].reshape(3, 3), cr.seg_pre[k][9:])
            v = q[s][cr.seg_qidx[k]]
            if cr.seg_type[k] == rd.JOINT_TRANS:
                T = T @ rd.frame(p=cr.seg_axis[k] * v)
            else:
                from scipy.spatial.transform import Rotation as Rot
                T = T @ rd.frame(Rot.from_rotvec(cr.seg_axis[k] * v).as_matrix())
            seg_T[k] = T.copy()
        by_link = {}
        for n, c, r in spheres:
            by_link.setdefault(n, []).append(np.array(c))
        want = []
        for i in range(cr.n_spheres):
            name = lt.names[cr.sph_link[i]]
            F = fr[cr.sph_link[i]]
            cands = [F[:9].reshape(3, 3) @ c + F[9:] for c in by_link[name]]
            got = seg_T[int(cr.sph_seg[i])][:3, :3] @ cr.sph_center[i] + seg_T[int(cr.sph_seg[i])][:3, 3]
            assert min(np.abs(got - c).max() for c in cands) < 1e-12


@pytest.mark.skipif(not os.path.exists("/root/reference"), reason="reference tree only exists in the build container")
def test_fixture_is_regenerable_from_reference(fixture_data):
    lt = fixture_data[0]
    cfg = "/root/reference/squirrel_8dof_planner/config/"
    lt2 = rd.link_table_from_urdf(open(cfg + "robotino_plan.urdf").read(), open(cfg + "robotino_plan.srdf").read())
    assert len(lt2.disabled_pairs) == 245
    rd.disable_rigid_pairs(lt2)
    assert lt2.names == lt.names
    assert np.array_equal(lt2.to_parent, lt.to_parent) and np.array_equal(lt2.parent, lt.parent)
    assert [tuple(p) for p in lt2.disabled_pairs] == [tuple(p) for p in lt.disabled_pairs]
    assert [(n, tuple(c), r) for n, c, r in sphere_model.spheres_for()] == fixture_data[2]


def test_world_and_state_generators_are_seeded():
    w1, w2 = syn.box_world(64, n_boxes=8, seed=1), syn.box_world(64, n_boxes=8, seed=1)
    assert np.array_equal(w1.bits, w2.bits) and w1.bits.dtype == np.uint32
    occ = w1.occupancy()
    assert occ[:8].all() and not occ[-1].all()            # floor slab below z = 0
    assert np.array_equal(syn.pack_occupancy(occ), w1.bits)
    s = syn.state_batch(1000, seed=42)
    assert s.shape == (1000, 8) and np.array_equal(s, syn.state_batch(1000, seed=42))
    a, b = syn.edge_batch(48, seed=42)
    assert np.array_equal(s[:21], syn.interpolate_edges(a[:1], b[:1]))
    d_rev = np.sqrt(((b - a)[:, 2:] ** 2).sum(1))
    d_pri = np.sqrt(((b - a)[:, :2] ** 2).sum(1))
    assert np.all(d_rev < 0.6 + 1e-12) and np.all(d_pri < 0.6 + 1e-12)     # one extend step (BS:5778-5799)


def test_shared_library_exports_the_declared_c_abi():
    hdr = open(os.path.join(ROOT, "include", "sqmp_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(sqmp_[a-z0-9_]+)\s*\(", hdr)))
    assert set(declared) == set(api.EXPORTS)
    L = api.load_library()
    for fn in declared:
        assert hasattr(L, fn), fn
    assert b"sm_100a" in L.sqmp_version()
    out = subprocess.run(["nm", "-D", "--defined-only", api.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (sqmp_[a-z0-9_]+)", out))
    assert set(declared) <= exported


def test_library_contains_only_sm100a_code():
    out = subprocess.run(["cuobjdump", "-lelf", api.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_gpu(fixture_data):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lt, fx, spheres = fixture_data
    with pytest.raises(api.SqmpError) as e:
        api.Context(rd.compile_robot(lt, spheres), 0)
    assert e.value.code == -5                                   # SQMP_ERR_NODEVICE
    from squirrel_motion_planner_b200 import planner_shim
    cc = planner_shim.CollisionChecker()
    assert not cc.initialized and cc.isInCollision(np.zeros(8), True, True)   # fails closed (CC:106-107)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "squirrel_motion_planner_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "sqmp_oracle" not in txt and "liboracle" not in txt, f


def test_shard_ranges_cover_and_align():
    for S in (1, 31, 32, 33, 1000, 1 << 20, (1 << 26) + 5):
        for R in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(R):
                lo, hi, per = sharding.shard_range(S, r, R)
                assert lo == prev or lo == S
                assert per % 32 == 0 and (lo % 32 == 0 or lo == S)
                prev = hi
            assert prev == S


WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
import oracle
from squirrel_motion_planner_b200 import robot_desc as rd, sharding, synthetic as syn
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=rank, world_size=world)
lt, fx = rd.load_fixture()
spheres = [(n, tuple(c), r) for n, c, r in fx["spheres"]]
orc = oracle.Oracle(lt, spheres)           # CPU stand-in for the kernel: the test is about the plumbing
w = syn.box_world(g=64, n_boxes=8, seed=5)
orc.set_grid(w.bits, w.dims, w.origin + np.array([0, 0, -0.02]), w.res)
S = 21 * 101 + 7
q = syn.state_batch(S, seed=42, xy_half=1.4)
lo, hi, per = sharding.shard_range(S, rank, world)
local = np.zeros(per // 32, dtype=np.uint32)
m = orc.check_states(q[lo:hi], oracle.FLAG_MAP)
local[:m.size] = m
full = sharding.gather_masks(torch.from_numpy(local.view(np.int32)), S, rank, world)
want = orc.check_states(q, oracle.FLAG_MAP)
assert np.array_equal(full.numpy().view(np.uint32), want), "gathered mask differs from the unsharded one"
assert 0 < sharding.unpack_mask(want, S).sum() < S
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_mask_allgather_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29600 + os.getpid() % 300)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT, port], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out


def test_sphere_model_is_conservative(fixture_data):
    """SURVEY.md §8 f3 / collision_checker.hpp:263-351: the shipped sphere model COVERS every shape the reference collides
    — the six URDF primitives (robotino_plan.urdf:323-330, 355-360, 385-390, 430-435, 546-551, 1369-1374) and the declared
    stand-ins of the 21 mesh links — and reaches at most one 0.05 m voxel outside it.  Coverage is proven, not sampled:
    every point of a lattice over the shape lies at least half a lattice diagonal inside some sphere."""
    from squirrel_motion_planner_b200 import sphere_fit as sf
    lt, fx, spheres = fixture_data
    shapes = sphere_model.link_shapes(lt)
    assert len(shapes) == 27                                          # the 27 shaped links of CC:263-351
    urdf = sorted(n for n, (_, src) in shapes.items() if src.startswith("URDF"))
    assert urdf == ["base_neck_link", "head_link", "hokuyo_link", "kinect_link", "neck_pan_link", "shell_base_link_front"]
    assert shapes["shell_base_link_front"][0][:2] == ("box", (0.55, 0.55, 0.17)) and shapes["neck_pan_link"][0][:2] == ("cylinder", (0.2, 0.08))
    rows = {r["link"]: r for r in sf.report(shapes, spheres, n=4000)}
    for name, row in rows.items():
        assert row["n_spheres"] >= 1                                  # every shaped link owns a sphere (disable mask, SRDF pairs)
        mine = [(c, r) for n, c, r in spheres if n == name]
        assert sf.verify_cover(row["shape"], [c for c, _ in mine], [r for _, r in mine]) == 0, name
        assert row["grade"]["covered"] >= 0.999, (name, row["grade"])
        assert row["grade"]["protrusion"] <= sphere_model.inflation_of(lt, name) + 1e-6 <= 0.05, (name, row["grade"])
    # the floor slab the planner inserts (squirrel_8dof_planner.cpp:889-902, top at z = -0.02 after CC:87) stays clear
    # of the lowest sphere: shapes start at z = +0.02 (CC:201), so their cover may reach at most 0.04 m below them
    assert sphere_model.P_BOTTOM < 0.04


def test_sphere_fit_tool(fixture_data):
    """The fitter itself: covers at other inflations, a thin box, a capsule chain; the grader reports a miss."""
    from squirrel_motion_planner_b200 import sphere_fit as sf
    for shape, p in ((("box", (0.3, 0.2, 0.1)), 0.03), (("cylinder", (0.1, 0.08)), 0.02), (("box", (0.035, 0.18, 0.02)), 0.035)):
        c, r = sf.cover_shape(shape, p)
        g = sf.grade(shape, c, r, n=4000)
        assert sf.verify_cover(shape, c, r) == 0 and g["covered"] == 1.0 and g["protrusion"] <= p + 1e-9
        small = r > r.min() + 1e-12
        assert not small.any() or sf.verify_cover(shape, c[small], r[small]) > 0   # without its smallest spheres (corners, rim) the cover has holes
    cap = ("capsule", ((0, 0, 0), (0, 0, 0.25), 0.05))
    c, r = sf.cover_shape(cap, 0.01)
    assert sf.verify_cover(cap, c, r) == 0 and abs(sf.protrusion(cap, c, r) - 0.01) < 1e-12
    assert sf.verify_cover(cap, c[::2], r[::2]) > 0                               # every other sphere: gaps on the surface
    assert sf.grade(cap, c, r, n=4000)["covered"] == 1.0
    assert sf.parse_shape("cylinder r=0.23 l=0.25")[:2] == ("cylinder", (0.23, 0.25)) and sf.parse_shape("mesh x.stl") is None
    assert np.allclose(sf.segment_distance([[0, 1, 0.5], [0, 0, 2]], [0, 0, 0], [0, 0, 1]), [1.0, 1.0])


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) needs no GPU: one JSON line with the
    contract's keys, the oracle port as `cpu_baseline`, zero-byte `e2e`."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "configs/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and line["gpu_launches"] == 0


def test_fp16_operand_pieces_of_the_tensor_core_scan_meet_their_stated_bounds():
    """csrc/tc_umma.cuh splits every coordinate into two fp16 pieces and |n|^2 into three (round to nearest, like
    __double2half); nn.cu's error budget for the tcgen05 estimate assumes |x - h - m| <= 2^-22 |x| + 2^-24 and
    | |n|^2 - p0 - p1 - p2 | <= 2^-30 |n|^2 + 2^-23 inside the range the kernel is used in (|x|^2 < 6e4).  numpy's
    float16 conversion rounds the same way, so the assumption is checked here without a GPU — including values whose
    second piece is an fp16 subnormal."""
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-244.0, 244.0, 200000), rng.uniform(-1, 1, 200000) * 10.0 ** rng.uniform(-9, 0, 200000),
                        [0.0, 6.0e-5, 6.2e-5, 244.9, -244.9, 1e-8, 5.96e-8]])
    h = x.astype(np.float16).astype(np.float64)
    m = (x - h).astype(np.float16).astype(np.float64)
    assert np.all(np.isfinite(h)) and np.all(np.abs(x - h - m) <= np.abs(x) * 2.0 ** -22 + 2.0 ** -24)
    assert np.all((-2.0 * h).astype(np.float16).astype(np.float64) == -2.0 * h)        # the image stores -2 h, -2 m exactly
    n2 = np.concatenate([rng.uniform(0.0, 6.0e4, 200000), 10.0 ** rng.uniform(-12, 0, 1000)])
    p0 = n2.astype(np.float16).astype(np.float64)
    p1 = (n2 - p0).astype(np.float16).astype(np.float64)
    p2 = (n2 - p0 - p1).astype(np.float16).astype(np.float64)
    assert np.all(np.isfinite(p0)) and np.all(np.abs(n2 - p0 - p1 - p2) <= n2 * 2.0 ** -30 + 2.0 ** -23)
