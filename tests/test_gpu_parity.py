"""GPU (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle and the committed goldens.

Bars (BASELINE.json north_star): near-neighbour indices bit-exact (ties -> lowest id); collision verdicts identical
except for states whose minimum clearance is within 1e-5 m of contact; FK within 1e-5 m / 1e-5 rad.
"""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, unpack
from squirrel_motion_planner_b200 import api, planner_shim, robot_desc as rd, synthetic as syn

pytestmark = pytest.mark.gpu
BAND = 1e-5          # metres — the north star's clearance band
FK_TOL = 1e-5        # metres / radians


def _verdict_parity(ctx, orc, q, flags, band=BAND):
    got = unpack(ctx.check_states(q, flags), len(q))
    ref_mask, cm, cs = orc.check_states(q, flags, clearance=True, nthreads=8)
    ref = unpack(ref_mask, len(q))
    clr = np.minimum(cm, cs)
    differs = got != ref
    assert not np.any(differs & (np.abs(clr) > band)), \
        "verdicts differ outside the band: %d states, |clearance| up to %g" % (differs.sum(), np.abs(clr[differs]).max())
    return got, ref, clr, differs


@pytest.fixture()
def world_ctx(ctx, orc, small_world):
    w, origin = small_world
    ctx.set_disabled_links(0)
    orc.set_disabled_links(0)
    ctx.set_voxel_grid(w.bits, w.dims, origin, w.res)
    orc.set_grid(w.bits, w.dims, origin, w.res)
    return ctx, orc, w, origin


# --------------------------------------------------------------------------------------- state validity
@pytest.mark.parametrize("flags", [api.CHECK_MAP, api.CHECK_SELF, api.CHECK_MAP | api.CHECK_SELF])
def test_check_states_matches_oracle(world_ctx, flags):
    ctx, orc, w, origin = world_ctx
    q = syn.state_batch(21 * 600 + 5, seed=17, xy_half=2.2)       # ragged: not a multiple of 32 or 128
    got, ref, clr, differs = _verdict_parity(ctx, orc, q, flags)
    assert 0.05 < got.mean() < 0.95
    if flags == api.CHECK_MAP:
        assert differs.sum() == 0                                 # fp64 exact stage: no disagreement at all here


def test_non_finite_states_fail_closed(world_ctx):
    """A joint value that is not finite (or absurd, |q| >= 1e7) makes the state "in collision" in both kernels, at both
    entry points, whatever the flags that check something — the oracle restates the same rule (spirit of CC:106-107)."""
    ctx, orc, w, origin = world_ctx
    q = syn.state_batch(256, seed=5, xy_half=0.3)
    bad = {3: np.nan, 40: np.inf, 41: -np.inf, 100: 1e30, 200: -2e7}
    for row, v in bad.items():
        q[row, row % 8] = v
    try:
        for kernel in (1, 2):
            ctx.set_option("k1_kernel", kernel)
            for flags in (api.CHECK_MAP, api.CHECK_SELF, api.CHECK_MAP | api.CHECK_SELF):
                got = unpack(ctx.check_states(q, flags), len(q))
                ref = unpack(orc.check_states(q, flags), len(q))
                assert got[list(bad)].all() and ref[list(bad)].all()
                assert np.array_equal(got, ref)
            assert not unpack(ctx.check_states(q, 0), len(q)).any()          # CC:109-110 comes first
        a = q[:64].copy()
        b = q[64:128].copy()
        b[5, 2] = np.nan                                                      # edge 5: every interpolated state but... all NaN
        m, first = ctx.check_edges(a, b, 20, api.CHECK_MAP)
        rm, rf = orc.check_edges(a, b, 20, oracle.FLAG_MAP)
        assert np.array_equal(m, rm) and np.array_equal(first, rf) and first[5] == 0
    finally:
        ctx.set_option("k1_kernel", 0)


def test_check_states_golden_numpy_world(ctx, orc):
    g = np.load(os.path.join(GOLDEN, "collision_numpy.npz"))
    ctx.set_voxel_grid(g["bits"], g["dims"], g["origin"], float(g["res"]))
    got = unpack(ctx.check_states(g["q"], api.CHECK_MAP), len(g["q"]))
    want = g["clearance"] <= 0.0
    assert not np.any((got != want) & (np.abs(g["clearance"]) > BAND))
    for s in range(0, len(g["q"]), 8):                                # getCollisions (CC:123-133)
        lm, _ = ctx.get_collisions(g["q"][s])
        assert np.array_equal(lm[:g["per_link"].shape[1]], g["per_link"][s])


def test_edge_cases_empty_flags_nogrid(fixture_data, small_world):
    lt, fx, spheres = fixture_data
    c2 = api.Context(rd.compile_robot(lt, spheres), 0)
    try:
        q = syn.state_batch(100, seed=3)
        assert not unpack(c2.check_states(q, api.CHECK_MAP), 100).any()          # CC:556-557: no octomap yet
        assert not unpack(c2.check_states(q, 0), 100).any()                        # CC:109-110
        assert c2.check_states(np.zeros((0, 8)), api.CHECK_MAP).size == 0           # empty batch
        w, origin = small_world
        c2.set_voxel_grid(w.bits, w.dims, origin, w.res)
        one = c2.check_states(q[:1], api.CHECK_MAP)
        assert one.size == 1
        # far outside the grid: unknown space is free
        far = q.copy(); far[:, 0] += 500.0
        assert not unpack(c2.check_states(far, api.CHECK_MAP), 100).any()
        with pytest.raises(api.SqmpError):
            c2.set_voxel_grid(w.bits, (0, 4, 4), origin, w.res)
    finally:
        c2.close()


@pytest.mark.parametrize("kernel", [1, 2])
def test_disabled_links_mask(world_ctx, fixture_data, kernel):
    ctx, orc, w, origin = world_ctx
    lt = fixture_data[0]
    q = syn.state_batch(21 * 200, seed=23, xy_half=2.2)
    names = ["base_body_link", "shell_base_link_front", "hand_wrist_link", "head_link"]
    mask = rd.link_mask(lt, names)
    try:
        ctx.set_option("k1_kernel", kernel)
        ctx.set_disabled_links(mask)
        orc.set_disabled_links(mask)
        got, ref, _, differs = _verdict_parity(ctx, orc, q, api.CHECK_MAP)
        assert differs.sum() == 0
        for one in ("neck_pan_link", "arm_link2", "shell_base_link"):      # single links, re-masked with no grid rebuild
            ctx.set_disabled_links(rd.link_mask(lt, [one]))
            orc.set_disabled_links(rd.link_mask(lt, [one]))
            assert _verdict_parity(ctx, orc, q[:1500], api.CHECK_MAP)[3].sum() == 0
        ctx.set_disabled_links(mask)
        ctx.set_disabled_links(0)
        full = unpack(ctx.check_states(q, api.CHECK_MAP), len(q))
        assert full.sum() > got.sum()                                  # disabling links can only free states
        ctx.set_disabled_links((1 << 64) - 1)
        assert not unpack(ctx.check_states(q, api.CHECK_MAP), len(q)).any()
    finally:
        ctx.set_option("k1_kernel", 0)
        ctx.set_disabled_links(0)
        orc.set_disabled_links(0)


def test_grid_partially_covering_the_robot(ctx, orc):
    """Spheres straddling the grid border and grids whose x size is not a multiple of 32."""
    w = syn.box_world(g=40, n_boxes=10, seed=8, dims=(37, 45, 21), z_low=-0.2)
    origin = w.origin + np.array([0.3, -0.2, rd.OCTREE_Z_OFFSET])
    ctx.set_voxel_grid(w.bits, w.dims, origin, w.res)
    orc.set_grid(w.bits, w.dims, origin, w.res)
    q = syn.state_batch(21 * 300, seed=29, xy_half=1.6)
    got, ref, _, differs = _verdict_parity(ctx, orc, q, api.CHECK_MAP)
    assert differs.sum() == 0 and 0.02 < got.mean() < 0.98


@pytest.mark.parametrize("res,dims", [(0.025, (150, 170, 90)), (0.08, (60, 45, 30)), (0.1, (40, 52, 25))])
def test_other_resolutions_match_oracle(ctx, orc, res, dims):
    """Grid resolution other than the reference's 0.05 m (sphere radii between 0.1 and 10 cells, other halo widths,
    other integer thresholds): both kernels against the oracle, no disagreement outside the band."""
    w = syn.box_world(g=64, res=res, n_boxes=14, seed=int(res * 1000), dims=dims, z_low=-0.3)
    origin = w.origin + np.array([0.07, -0.11, rd.OCTREE_Z_OFFSET])
    ctx.set_disabled_links(0)
    orc.set_disabled_links(0)
    ctx.set_voxel_grid(w.bits, w.dims, origin, w.res)
    orc.set_grid(w.bits, w.dims, origin, w.res)
    q = syn.state_batch(21 * 250, seed=97, xy_half=min(dims[0], dims[1]) * res / 2.0 * 1.2)
    try:
        for kernel in (1, 2):
            ctx.set_option("k1_kernel", kernel)
            got, ref, clr, differs = _verdict_parity(ctx, orc, q, api.CHECK_MAP)
            assert differs.sum() == 0 and 0.02 < got.mean() < 0.995   # small cluttered worlds: most states collide
    finally:
        ctx.set_option("k1_kernel", 0)


def test_real_scene_room3_with_floor_patch(ctx, orc):
    """The reference's own room3.bt octomap (tests/golden/rooms_bt.npz), run through the .bt writer/parser, with the
    planner node's floor patch (squirrel_8dof_planner.cpp:886-904) around the robot: a 202 x 322 x 15 grid (not a cube,
    not a power of two), thin obstacles, most of the volume unknown = free."""
    from squirrel_motion_planner_b200 import octomap_io as oio
    g = np.load(os.path.join(GOLDEN, "rooms_bt.npz"))
    w0 = syn.VoxelWorld(g["room3_bits"], tuple(int(v) for v in g["room3_dims"]), g["room3_origin"], float(g["room3_res"]))
    res, occ, _free, _n = oio.parse_bt(oio.write_bt(w0))
    w = oio.to_voxel_world(res, oio.insert_floor(occ, res, (0.5, -1.0), 3.0))
    origin = w.origin + np.array([0.0, 0.0, rd.OCTREE_Z_OFFSET])
    ctx.set_disabled_links(0)
    orc.set_disabled_links(0)
    ctx.set_voxel_grid(w.bits, w.dims, origin, w.res)
    orc.set_grid(w.bits, w.dims, origin, w.res)
    q = syn.state_batch(21 * 400, seed=29, xy_half=4.5)
    q[:, 1] *= 1.6                                                      # the room is 10 m x 16 m
    for flags in (api.CHECK_MAP, api.CHECK_MAP | api.CHECK_SELF):
        got, ref, clr, differs = _verdict_parity(ctx, orc, q, flags)
        assert differs.sum() == 0
    assert 0.02 < got.mean() < 0.98


# --------------------------------------------------------------------------------------- K5: batched sampling (f4)
@pytest.mark.parametrize("specialized", [1, 0])
def test_sample_configs_bit_exact_and_order_free(ctx, orc, specialized):
    """sqmp_sample_configs == the oracle's counter stream bit for bit (configs and try counts); acceptance decided by
    the kernel's own FK may differ from the CPU's only if |ee_z| < 1e-9 (never observed)."""
    lo = np.array([-5.0, -4.0, -np.pi, -1.2, -1.7, -1.8, -2.4, -2.9])
    hi = np.array([5.0, 6.0, np.pi, 1.5, 2.6, 1.8, 2.4, 2.9])
    ctx.set_option("chain_specialization", specialized)
    try:
        for seed, first, n, T in ((1, 0, 5000, 64), (0xFEEDFACE12345, 123456789, 1031, 3)):
            q, tries = ctx.sample_configs(seed, first, n, lo, hi, T)
            qr, tr = orc.sample_configs(seed, first, n, lo, hi, T)
            same = np.all(q == qr, axis=1) & (tries == tr)
            if not same.all():
                assert np.abs(orc.fk(qr[~same])[:, 2]).max() < 1e-9
            assert same.mean() > 0.999
            assert np.all(q >= lo) and np.all(q <= hi)
            ee = ctx.fk(q)
            assert np.all((tries > 0) == (ee[:, 2] >= 0.0))
        a, _ = ctx.sample_configs(7, 40, 300, lo, hi, 8)
        b, _ = ctx.sample_configs(7, 0, 340, lo, hi, 8)
        assert np.array_equal(a, b[40:])
        # rejection-heavy box (arm folded under the platform): exhausted tries are reported as -max_tries
        lo2, hi2 = lo.copy(), hi.copy()
        lo2[3:] = [1.3, -1.7, -1.8, -0.4, -2.9]; hi2[3:] = [1.5, -1.0, -0.8, 0.8, 2.9]
        q2, t2 = ctx.sample_configs(3, 0, 2000, lo2, hi2, 4)
        qr2, tr2 = orc.sample_configs(3, 0, 2000, lo2, hi2, 4)
        assert np.array_equal(q2, qr2) and np.array_equal(t2, tr2)
        assert (t2 > 1).any() and (t2 < 0).any()
        assert ctx.sample_configs(1, 0, 0, lo, hi, 4)[0].shape == (0, 8)            # empty batch
        with pytest.raises(api.SqmpError):
            ctx.sample_configs(1, 0, 4, hi, lo, 4)                                    # lo > hi
        with pytest.raises(api.SqmpError):
            ctx.sample_configs(1, 0, 4, lo, hi, 0)                                    # no tries
    finally:
        ctx.set_option("chain_specialization", 1)


@pytest.mark.parametrize("specialized", [1, 0])
def test_informed_sampler_bit_exact(ctx, orc, specialized):
    """sqmp_sample_informed (sampleJointConfigfromEllipse_JntArray, BS:3607-3829) == the oracle's restatement, double for
    double: same counter stream, same operation order; narrow and wide ellipses, a confined and an unconfined environment."""
    from squirrel_motion_planner_b200 import informed
    lo = np.array([-5.0, -4.0, -np.pi, -1.2, -1.7, -1.8, -2.4, -2.9])
    hi = np.array([5.0, 6.0, np.pi, 1.5, 2.6, 1.8, 2.4, 2.9])
    start = np.array([-2.0, -1.5, 0.3, -0.17, 1.96, 0.19, 0.65, 0.53])
    goal = np.array([1.5, 1.0, 1.2, -0.2, 1.2, 0.4, 0.9, 0.3])
    el = informed.ellipse_init(start, goal, syn.IS_REVOLUTE)
    ctx.set_option("chain_specialization", specialized)
    try:
        for k_rev, k_pri, env in ((1.0001, 1.0001, True), (1.5, 1.2, True), (4.0, 3.0, True), (2.0, 2.0, False)):
            es = informed.ctypes_struct(el, k_rev * el["rev"]["c_min"], k_pri * el["pri"]["c_min"],
                                        (lo[0], hi[0]) if env else (0.0, 0.0), (lo[1], hi[1]) if env else (0.0, 0.0))
            for seed, first, n, T in ((3, 0, 4000, 64), (0xC0FFEE1234, 987654321, 777, 2)):
                q, t = ctx.sample_informed(seed, first, n, lo, hi, es, T)
                qr, tr = orc.sample_informed(seed, first, n, lo, hi, es, T)
                same = np.all(q == qr, axis=1) & (t == tr)
                if not same.all():                                  # only where the acceptance test sits on its boundary
                    assert np.abs(orc.fk(qr[~same])[:, 2]).max() < 1e-9
                assert same.mean() > 0.999
        a, _ = ctx.sample_informed(9, 10, 90, lo, hi, es, 8)
        b, _ = ctx.sample_informed(9, 0, 100, lo, hi, es, 8)
        assert np.array_equal(a, b[10:])
        bad = informed.ctypes_struct(el, 3.0, 3.0)
        bad.n_rev = 7
        with pytest.raises(api.SqmpError):
            ctx.sample_informed(1, 0, 4, lo, hi, bad, 4)
    finally:
        ctx.set_option("chain_specialization", 1)


# --------------------------------------------------------------------------------------- the two validity kernels
@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("flags", [api.CHECK_MAP, api.CHECK_MAP | api.CHECK_SELF])
def test_both_validity_kernels_match_oracle(world_ctx, kernel, flags):
    """k1_kernel = 1: generic fp64 kernel (validity.cu); 2: fast kernel (validity_fast.cu: fp32 proof stages with an
    error bound on halo-padded fields, fp64 exact stage).  Both must reproduce the oracle with no disagreement."""
    ctx, orc, w, origin = world_ctx
    ctx.set_option("k1_kernel", kernel)
    try:
        q = syn.state_batch(21 * 500 + 11, seed=41, xy_half=2.2)
        got, ref, clr, differs = _verdict_parity(ctx, orc, q, flags)
        assert differs.sum() == 0 and 0.05 < got.mean() < 0.95
    finally:
        ctx.set_option("k1_kernel", 0)


def test_fast_kernel_is_bit_identical_to_generic_kernel(ctx, orc, fixture_data):
    """Same masks bit for bit — including states whose base lies outside the grid (the fast kernel's forced lanes),
    spheres straddling every face of the grid (scan boxes clamped to odd row counts), grids that are not cubes, a
    disabled-link mask, the edge entry point and the self-collision instantiation."""
    lt = fixture_data[0]
    worlds = [
        (syn.box_world(g=96, n_boxes=20, seed=5, dims=(96, 80, 64)), np.zeros(3), 2.2),
        (syn.box_world(g=40, n_boxes=10, seed=8, dims=(37, 45, 21), z_low=-0.2), np.array([0.3, -0.2, 0.0]), 1.6),
        (syn.box_world(g=64, n_boxes=60, seed=77, dims=(33, 64, 70)), np.array([-0.4, 0.1, 0.0]), 1.4),
    ]
    try:
        for w, shift, half in worlds:
            origin = w.origin + shift + np.array([0.0, 0.0, rd.OCTREE_Z_OFFSET])
            ctx.set_voxel_grid(w.bits, w.dims, origin, w.res)
            orc.set_grid(w.bits, w.dims, origin, w.res)
            for wide in (1.0, 1.7, 3.0):                                   # 1.7, 3.0: many bases outside the grid
                q = syn.state_batch(21 * 700 + 3, seed=int(10 * wide) + 51, xy_half=half * wide)
                for mask in (0, rd.link_mask(lt, ["head_link", "base_body_link"])):
                    ctx.set_disabled_links(mask)
                    for flags in (api.CHECK_MAP, api.CHECK_MAP | api.CHECK_SELF):
                        ctx.set_option("k1_kernel", 1)
                        a = ctx.check_states(q, flags)
                        ctx.set_option("k1_kernel", 2)
                        b = ctx.check_states(q, flags)
                        assert np.array_equal(a, b)
                ctx.set_disabled_links(0)
                got = unpack(b, len(q))
                assert 0.01 < got.mean() < 0.99
                if wide == 1.7:                                            # and the oracle agrees on the wide batch
                    orc.set_disabled_links(0)
                    _verdict_parity(ctx, orc, q[:4000], api.CHECK_MAP)
            ea, eb = syn.edge_batch(333, seed=61, xy_half=half * 1.5)
            ctx.set_option("k1_kernel", 1)
            m1, f1 = ctx.check_edges(ea, eb, 20, api.CHECK_MAP)
            ctx.set_option("k1_kernel", 2)
            m2, f2 = ctx.check_edges(ea, eb, 20, api.CHECK_MAP)
            assert np.array_equal(m1, m2) and np.array_equal(f1, f2)
    finally:
        ctx.set_option("k1_kernel", 0)
        ctx.set_disabled_links(0)


def test_fast_kernel_chain_interpreter_equals_compile_time_chain(world_ctx):
    """K1F walks the chain through the compile-time signature of the reference's robot when the description matches it
    and through the per-segment interpreter otherwise: same frames, same masks."""
    ctx, orc, w, origin = world_ctx
    q = syn.state_batch(21 * 900 + 7, seed=71, xy_half=2.4)
    ctx.set_option("k1_kernel", 2)
    try:
        for flags in (api.CHECK_MAP, api.CHECK_MAP | api.CHECK_SELF):
            ctx.set_option("chain_specialization", 1)
            a = ctx.check_states(q, flags)
            ctx.set_option("chain_specialization", 0)
            b = ctx.check_states(q, flags)
            assert np.array_equal(a, b)
    finally:
        ctx.set_option("chain_specialization", 1)
        ctx.set_option("k1_kernel", 0)


def test_back_to_back_launches_keep_serial_semantics(world_ctx):
    """K1F grids are launched with programmatic stream serialization: a grid may start while its predecessor drains,
    but it writes its mask only after the predecessor has completed.  Calls that reuse ONE mask buffer with no
    synchronisation in between must therefore leave the result of the last call — for one-wave and multi-wave grids."""
    import torch
    ctx, orc, w, origin = world_ctx
    stream = torch.cuda.current_stream().cuda_stream
    for S in (1 << 11, 1 << 15, 1 << 18):
        d_q = [torch.from_numpy(syn.state_batch(S, seed=200 + i, xy_half=2.2)).cuda() for i in range(4)]
        d_m = torch.zeros(S // 32, dtype=torch.int32, device="cuda")
        want = [ctx.check_states(q.cpu().numpy(), api.CHECK_MAP).view(np.int32) for q in d_q]
        for overlap in (1, 0):
            ctx.set_option("launch_overlap", overlap)
            for trial in range(6):
                order = np.random.default_rng(trial).permutation(4)
                for i in order:
                    ctx.check_states_dev(d_q[i], S, api.CHECK_MAP, d_m, stream)
                torch.cuda.synchronize()
                assert np.array_equal(d_m.cpu().numpy(), want[order[-1]])
    ctx.set_option("launch_overlap", 0)


def test_shared_voxel_grid_between_contexts(world_ctx, fixture_data):
    """sqmp_share_voxel_grid: a second context on the same device borrows the owner's fields (one copy in HBM) and
    returns the same masks; its own disable mask and a later grid of its own do not disturb the owner."""
    ctx, orc, w, origin = world_ctx
    lt, fx, spheres = fixture_data
    q = syn.state_batch(21 * 300, seed=81, xy_half=2.2)
    want = ctx.check_states(q, api.CHECK_MAP | api.CHECK_SELF)
    c2 = api.Context(rd.compile_robot(lt, spheres), 0)
    try:
        with pytest.raises(api.SqmpError):
            ctx.share_voxel_grid(ctx)                                      # a context cannot borrow from itself
        c2.share_voxel_grid(ctx)
        assert np.array_equal(c2.check_states(q, api.CHECK_MAP | api.CHECK_SELF), want)
        c2.set_disabled_links(rd.link_mask(lt, ["head_link"]))             # per-context, the fields stay shared
        assert np.array_equal(ctx.check_states(q, api.CHECK_MAP | api.CHECK_SELF), want)
        w2 = syn.box_world(g=48, n_boxes=4, seed=2)
        c2.set_voxel_grid(w2.bits, w2.dims, w2.origin, w2.res)            # an own grid replaces the borrowed one
        assert np.array_equal(ctx.check_states(q, api.CHECK_MAP | api.CHECK_SELF), want)
    finally:
        c2.close()
    assert np.array_equal(ctx.check_states(q, api.CHECK_MAP | api.CHECK_SELF), want)   # the owner outlives the borrower


def test_fused_gather_entry_point_single_rank(world_ctx):
    """sqmp_check_states_gather_dev / sqmp_gather_signal_dev / sqmp_gather_wait_dev with one "peer" (this GPU itself): the
    words land at word_offset of the gather buffer, the flag shows the step, the wait kernel reports success; a step
    that was never signalled makes the wait time out (status 1) instead of hanging.  (The 2- and 8-GPU runs of bench.py
    --gather p2p check the peer-mapped case against ncclAllGather.)"""
    import ctypes as C
    import torch
    ctx, orc, w, origin = world_ctx
    stream = torch.cuda.current_stream().cuda_stream
    S = 21 * 1000 + 13
    q = syn.state_batch(S, seed=91, xy_half=2.2)
    d_q = torch.from_numpy(q).cuda()
    words = (S + 31) // 32
    off = 17
    d_buf = torch.zeros(off + words + 5, dtype=torch.int32, device="cuda")
    d_flags = torch.zeros(8, dtype=torch.int32, device="cuda")
    d_status = torch.full((1,), -1, dtype=torch.int32, device="cuda")
    g = api.PeerGatherC()
    g.n_peers, g.rank, g.word_offset = 1, 0, off
    g.mask[0], g.flag[0] = d_buf.data_ptr(), d_flags.data_ptr()
    ctx.check_states_gather_dev(d_q, S, api.CHECK_MAP, g, stream)
    ctx.gather_signal_dev(g, 5, stream)
    ctx.gather_wait_dev(d_flags.data_ptr(), 1, 5, d_status, stream)
    torch.cuda.synchronize()
    want = ctx.check_states(q, api.CHECK_MAP).view(np.int32)
    got = d_buf.cpu().numpy()
    assert np.array_equal(got[off:off + words], want) and not got[:off].any() and not got[off + words:].any()
    assert int(d_flags[0].item()) == 5 and int(d_status.item()) == 0
    ctx.gather_wait_dev(d_flags.data_ptr(), 1, 6, d_status, stream)          # step 6 is never signalled
    torch.cuda.synchronize()
    assert int(d_status.item()) == 1
    with pytest.raises(api.SqmpError):
        g.n_peers = 9
        ctx.check_states_gather_dev(d_q, S, api.CHECK_MAP, g, stream)


def test_fast_kernel_rejected_without_its_preconditions(fixture_data):
    """k1_kernel = 2 is an error for a description whose base is not a translation-only run (no silent fallback)."""
    lt, fx, spheres = fixture_data
    cr = rd.compile_robot(lt, spheres)
    c2 = api.Context(cr, 0)
    try:
        c2.set_option("k1_kernel", 2)                                    # the reference robot qualifies
        c2.set_option("k1_kernel", 0)
        with pytest.raises(api.SqmpError):
            c2.set_option("k1_kernel", 3)
    finally:
        c2.close()


def test_check_edges_matches_oracle_and_states_entry(world_ctx):
    ctx, orc, w, origin = world_ctx
    a, b = syn.edge_batch(777, seed=31, xy_half=2.2)
    mask, first = ctx.check_edges(a, b, 20, api.CHECK_MAP)
    ref_mask, ref_first = orc.check_edges(a, b, 20, oracle.FLAG_MAP, nthreads=8)
    assert np.array_equal(mask, ref_mask) and np.array_equal(first, ref_first)
    # device-side interpolation == host-side interpolation (BS:4461-4478), state for state
    states = syn.interpolate_edges(a, b, 20)
    assert np.array_equal(ctx.check_states(states, api.CHECK_MAP), mask)
    assert (first == -1).any() and (first == 0).any() and (first > 1).any()
    m5, f5 = ctx.check_edges(a[:50], b[:50], 7, api.CHECK_MAP | api.CHECK_SELF)
    r5, rf5 = orc.check_edges(a[:50], b[:50], 7, oracle.FLAG_MAP | oracle.FLAG_SELF)
    assert np.array_equal(m5, r5) and np.array_equal(f5, rf5)


def test_shim_collision_checker_and_edge_validity(fixture_data, small_world, orc):
    w, origin = small_world
    cc = planner_shim.CollisionChecker(0)
    assert cc.initialized
    cc.setOcTree(None)                                                # CC:78-79
    cc.setOcTree(w)                                                   # applies the -0.02 offset itself
    orc.set_grid(w.bits, w.dims, origin, w.res)
    hp = planner_shim.BiRRTstarHotPath(cc)
    q = syn.state_batch(21 * 40, seed=37, xy_half=2.2)
    for s in range(0, len(q), 17):
        for cs_, cm_ in ((True, True), (False, True), (True, False), (False, False)):
            assert cc.isInCollision(q[s], cs_, cm_) == orc.is_in_collision(q[s], cs_, cm_)
            assert hp.isConfigValid(q[s], cs_, cm_) == (not orc.is_in_collision(q[s], cs_, cm_))
    for e in range(40):
        traj = q[e * 21:(e + 1) * 21]
        assert hp.isEdgeValid(planner_shim.Edge(traj), False, True) == orc.is_edge_valid(traj, False, True)
        assert hp.isEdgeValid(planner_shim.Edge(traj), True, True) == orc.is_edge_valid(traj, True, True)
    pairs, links = cc.getCollisions(q[3])
    lm, ph = orc.get_collisions(q[3])
    lt = fixture_data[0]
    assert links == sorted(lt.names[i] for i in np.nonzero(lm)[0])
    want_pairs = set((lt.names[i], lt.names[j]) for i, j in zip(*np.nonzero(ph)) if i < j)
    # the kernel reports pairs across rigid groups; pairs inside one rigid group are out of the matrix (robot_desc.disable_rigid_pairs)
    assert set(pairs) == want_pairs
    for s_ in range(0, len(q), 29):                                    # getCollisions over more states, both outputs
        pairs, links = cc.getCollisions(q[s_])
        lm, ph = orc.get_collisions(q[s_])
        assert links == sorted(lt.names[i] for i in np.nonzero(lm)[0])
        assert set(pairs) == set((lt.names[i], lt.names[j]) for i, j in zip(*np.nonzero(ph)) if i < j)
    cc.setDisabledLinkMapCollisions(["base_body_link"])
    orc.set_disabled_links(rd.link_mask(lt, ["base_body_link"]))
    assert cc.isInCollision(q[3], False, True) == orc.is_in_collision(q[3], False, True)
    orc.set_disabled_links(0)


# --------------------------------------------------------------------------------------- FK
def _fk_err(a, b):
    pos = np.abs(a[:, :3] - b[:, :3]).max()
    # rotation angle between the two orientations: 2*asin(|dq|/2) ~ |dq| * 2 for unit quaternions aligned in sign
    # (2*acos(dot) is ill-conditioned near 1 and would report sqrt(eps) for identical rotations)
    sgn = np.sign((a[:, 3:] * b[:, 3:]).sum(1))[:, None]
    ang = (2.0 * np.arcsin(np.clip(np.linalg.norm(a[:, 3:] - sgn * b[:, 3:], axis=1) / 2.0, 0, 1))).max()
    return pos, ang


@pytest.mark.parametrize("specialized", [1, 0])
@pytest.mark.parametrize("prec", [64, 32])
def test_fk_matches_oracle_and_golden(ctx, orc, prec, specialized):
    """specialized=1: compile-time unrolled chain (signature of the reference's robot); 0: generic chain interpreter."""
    ctx.set_option("fk_precision", prec)
    ctx.set_option("chain_specialization", specialized)
    try:
        g = np.load(os.path.join(GOLDEN, "fk_scipy.npz"))
        got = ctx.fk(g["q"])
        pos, ang = _fk_err(got, g["ee"])
        assert pos < FK_TOL and ang < FK_TOL
        for n in (1, 127, 128, 129, 1000, 4097):                       # ragged tiles around the 128-state CTA tile
            q = syn.sample_configs(n, seed=41 + n, xy_half=12.0)
            got, ref = ctx.fk(q), orc.fk(q)
            pos, ang = _fk_err(got, ref)
            assert pos < FK_TOL and ang < FK_TOL, (n, pos, ang)
            if prec == 64:
                assert np.abs(got - ref).max() < 1e-12               # same quaternion sign convention too
    finally:
        ctx.set_option("fk_precision", 64)
        ctx.set_option("chain_specialization", 1)


# --------------------------------------------------------------------------------------- near neighbours
@pytest.mark.parametrize("variant", [1, 2, 3, 4])
def test_nearest_bit_exact(ctx, variant):
    ctx.set_option("nn_variant", variant)
    try:
        g = np.load(os.path.join(GOLDEN, "nn_numpy.npz"))
        ctx.tree_reset(0)
        assert ctx.tree_append_batch(0, g["nodes"], g["cost"]) == 0
        assert np.array_equal(ctx.nn_nearest(0, g["queries"]), g["nearest"])
        for N, Q, seed in ((1, 5, 1), (31, 40, 2), (257, 129, 3), (20011, 300, 4)):
            nodes, cost = syn.random_tree(N, seed=seed, cost_seed=seed + 1)
            if N > 100:
                nodes[N // 2] = nodes[7]; nodes[N - 1] = nodes[7]      # duplicates -> ties
            queries = syn.sample_configs(Q, seed=seed + 2)
            queries[0] = nodes[min(7, N - 1)]
            ctx.tree_reset(1)
            ctx.tree_append_batch(1, nodes, cost)
            assert np.array_equal(ctx.nn_nearest(1, queries), oracle.nearest(nodes, queries, nthreads=8)), (N, Q)
    finally:
        ctx.set_option("nn_variant", 0)


@pytest.mark.parametrize("variant", [1, 2, 3, 4])
def test_nn_against_reference_golden(ctx, variant):
    """tests/golden/nn_reference.npz = outputs of the reference's own DistanceHeuristics + Node::operator< / std::sort
    (compiled from /root/reference by oracle/Makefile.ref): ids bit-exact; within a group of equal cost the reference's
    unstable sort may order ids either way, the CUDA path orders them by id."""
    g = np.load(os.path.join(GOLDEN, "nn_reference.npz"))
    nodes, cost, queries = g["nodes"], g["cost"], g["queries"]
    ctx.set_option("nn_variant", variant)
    try:
        ctx.tree_reset(4)
        ctx.tree_append_batch(4, nodes, cost)
        assert np.array_equal(ctx.nn_nearest(4, queries), g["nearest"])
        low, high, cnt = ctx.nn_radius(4, queries, None, 1.5, 20)
        assert np.array_equal(cnt, g["near_count"])
        off = 0
        for k in range(len(queries)):
            n = int(cnt[k]); want = g["near_flat"][off:off + n]; off += n
            got_lo, got_hi = low[k][low[k] >= 0], high[k][high[k] >= 0]
            assert np.array_equal(cost[got_lo], cost[want[:20]]) and np.array_equal(cost[got_hi], cost[want[::-1][:20]])
            if n <= 20:
                assert sorted(got_lo.tolist()) == sorted(want.tolist())
    finally:
        ctx.set_option("nn_variant", 0)


def test_nearest_quirks_and_tree_mutation(ctx):
    ctx.tree_reset(2)
    assert ctx.nn_nearest(2, np.zeros((3, 8))).tolist() == [0, 0, 0]   # empty tree: id stays 0 (BS:4084)
    far = np.full((5, 8), 1e6)
    ctx.tree_append_batch(2, far, np.zeros(5))
    assert ctx.nn_nearest(2, np.zeros((1, 8)))[0] == 0                  # nothing below 10000.0 (BS:4090)
    nid = ctx.tree_append(2, np.full(8, 2.0), 1.0)
    assert nid == 5 and ctx.tree_size(2) == 6
    assert ctx.nn_nearest(2, np.zeros((1, 8)))[0] == 5
    ctx.tree_update(2, 1, np.full(8, 1.0), 2.0)                         # rewire moves a node (BS:5173-5193)
    assert ctx.nn_nearest(2, np.zeros((1, 8)))[0] == 1
    for i in range(300):                                                # growth past the initial capacity
        ctx.tree_append(2, np.full(8, 50.0 + i), float(i))
    assert ctx.tree_size(2) == 306 and ctx.nn_nearest(2, np.full((1, 8), 60.2))[0] == 16
    ctx.tree_reset(2)
    assert ctx.tree_size(2) == 0


def test_deferred_tree_writes_keep_program_order(ctx):
    """Single-node appends / updates wait in a host-side queue and reach the device in one launch before the next call
    that reads a tree (capi.cu: queue_tree_write).  Two trees interleaved, nodes updated several times before a read, an
    update right after its append, more writes than the queue holds, a batch append and a cost update in between: every
    answer equals the oracle's on the tree the program order defines."""
    rng = np.random.default_rng(31)
    mirror = {3: ([], []), 4: ([], [])}
    for tr in (3, 4):
        ctx.tree_reset(tr)
    def check():
        for tr in (3, 4):
            nodes, cost = np.array(mirror[tr][0]), np.array(mirror[tr][1])
            qs = np.concatenate([nodes[rng.integers(0, len(nodes), 8)] + 1e-3, syn.sample_configs(8, seed=int(rng.integers(1 << 30)))])
            assert np.array_equal(ctx.nn_nearest(tr, qs), oracle.nearest(nodes, qs, nthreads=4))
            got = ctx.nn_radius(tr, qs, None, 2.0, 5)
            ref = oracle.near_topk(nodes, cost, qs, None, 2.0, 5, nthreads=4)
            for a, b in zip(got, ref):
                assert np.array_equal(a, b)
    for step in range(700):
        tr = 3 + int(rng.integers(0, 2))
        nodes, cost = mirror[tr]
        kind = rng.integers(0, 10)
        if kind < 6 or len(nodes) < 4:
            q, c = syn.sample_configs(1, seed=int(rng.integers(1 << 30)))[0], float(rng.uniform(0, 5))
            assert ctx.tree_append(tr, q, c) == len(nodes)
            nodes.append(q); cost.append(c)
            if kind == 0:                                               # rewired at once
                nodes[-1] = q + 0.5; cost[-1] = c + 1.0
                ctx.tree_update(tr, len(nodes) - 1, nodes[-1], cost[-1])
        elif kind < 9:
            i = int(rng.integers(0, len(nodes)))
            for _ in range(int(rng.integers(1, 4))):                    # the last write wins
                nodes[i] = syn.sample_configs(1, seed=int(rng.integers(1 << 30)))[0]; cost[i] = float(rng.uniform(0, 5))
                ctx.tree_update(tr, i, nodes[i], cost[i])
        else:
            i = int(rng.integers(0, len(nodes)))
            cost[i] = float(rng.uniform(0, 5))
            ctx.tree_update_costs(tr, np.array([i], dtype=np.int32), np.array([cost[i]]))
        if step in (5, 40, 41, 400, 699):                               # (400: well past the 256 records of the queue)
            check()
        if step == 200:
            extra = syn.sample_configs(10, seed=77)
            ctx.tree_append_batch(tr, extra, np.ones(10))
            nodes.extend(list(extra)); cost.extend([1.0] * 10)


@pytest.mark.parametrize("variant", [1, 2, 3, 4])
def test_radius_topk_bit_exact(ctx, variant):
    ctx.set_option("nn_variant", variant)
    try:
        g = np.load(os.path.join(GOLDEN, "nn_numpy.npz"))
        ctx.tree_reset(0)
        ctx.tree_append_batch(0, g["nodes"], g["cost"])
        low, high, cnt = ctx.nn_radius(0, g["queries"], None, 1.5, 20)
        assert np.array_equal(cnt, g["near_count"])
        off = 0
        for k in range(len(g["queries"])):
            n = int(cnt[k]); want = g["near_flat"][off:off + n]; off += n
            assert np.array_equal(low[k][low[k] >= 0], want[:20])
            assert np.array_equal(high[k][high[k] >= 0], want[::-1][:20])
        for N, Q, r, k, seed in ((500, 33, 3.0, 20, 11), (20011, 200, 2.2, 20, 12), (3000, 64, 50.0, 7, 13)):
            nodes, cost = syn.random_tree(N, seed=seed, cost_seed=seed + 1)
            cost[::7] = cost[3]                                        # equal costs -> id tie-break
            queries = syn.sample_configs(Q, seed=seed + 2)
            excl = np.random.default_rng(seed).integers(-1, N, size=Q).astype(np.int32)
            ctx.tree_reset(1)
            ctx.tree_append_batch(1, nodes, cost)
            got = ctx.nn_radius(1, queries, excl, r, k)
            want = oracle.near_topk(nodes, cost, queries, excl, r, k, nthreads=8)
            for a, b, name in zip(got, want, ("low", "high", "count")):
                assert np.array_equal(a, b), (N, Q, r, name)
            # cost updates (recursiveNodeCostUpdate, BS:5518-5520)
            ids = np.arange(0, N, 5, dtype=np.int32)
            cost[ids] = cost[ids] * 0.5 + 1.0
            ctx.tree_update_costs(1, ids, cost[ids])
            got = ctx.nn_radius(1, queries, excl, r, k)
            want = oracle.near_topk(nodes, cost, queries, excl, r, k, nthreads=8)
            for a, b in zip(got, want):
                assert np.array_equal(a, b)
    finally:
        ctx.set_option("nn_variant", 0)


def test_multi_tree_queries_equal_single_tree_calls(ctx):
    """sqmp_nn_nearest_multi / sqmp_nn_radius_multi (query i against tree[i], one launch) == the oracle on every tree:
    trees of very different sizes incl. an empty one and a one-node tree, duplicated nodes (ties -> lowest id), equal
    costs (ties -> lowest id), exclusions, more hits than k, and a pool small enough to force the overflow fall-back."""
    rng = np.random.default_rng(3)
    sizes = [0, 1, 7, 33, 500, 3000, 64, 1025]
    trees, costs = [], []
    for t, n in enumerate(sizes):
        ctx.tree_reset(10 + t)
        nodes, cost = syn.random_tree(max(n, 1), seed=300 + t, cost_seed=400 + t, xy_half=1.0)
        nodes, cost = nodes[:n], cost[:n]
        if n >= 33:
            nodes[5] = nodes[20]; nodes[31] = nodes[20]            # exact duplicates
            cost[::4] = cost[3]                                     # equal costs
        if n:
            ctx.tree_append_batch(10 + t, nodes, cost)
        trees.append(nodes); costs.append(cost)
    Q = 400
    which = rng.integers(0, len(sizes), Q)
    queries = syn.sample_configs(Q, seed=77, xy_half=1.0)
    for i in range(0, Q, 9):                                        # some queries ARE nodes (distance 0, exclusion)
        if sizes[which[i]] >= 33:
            queries[i] = trees[which[i]][20]
    excl = np.where(rng.random(Q) < 0.5, 20, -1).astype(np.int32)
    ids = ctx.nn_nearest_multi(10 + which, queries)
    for i in range(Q):
        n = sizes[which[i]]
        want = 0 if n == 0 else int(oracle.nearest(trees[which[i]], queries[i:i + 1])[0])
        assert ids[i] == want
    for pool in (1 << 30, 1 << 16):                                 # 2^16 bytes / 400 queries: 40-hit slices overflow
        ctx.set_option("nn_hit_pool_bytes", pool)
        low, high, cnt = ctx.nn_radius_multi(10 + which, queries, excl, 3.0, 20)
        for i in range(Q):
            n = sizes[which[i]]
            if n == 0:
                assert cnt[i] == 0 and (low[i] == -1).all() and (high[i] == -1).all()
                continue
            rl, rh, rc = oracle.near_topk(trees[which[i]], costs[which[i]], queries[i:i + 1], excl[i:i + 1], 3.0, 20)
            assert cnt[i] == rc[0] and np.array_equal(low[i], rl[0]) and np.array_equal(high[i], rh[0])
    ctx.set_option("nn_hit_pool_bytes", 1 << 30)
    assert cnt.max() > 40                                           # the small pool really overflowed somewhere
    with pytest.raises(api.SqmpError):
        ctx.nn_nearest_multi(np.array([5000], dtype=np.int32), queries[:1])


@pytest.mark.parametrize("variant", [3, 4])
def test_fp32_prefilter_is_sound_on_adversarial_trees(ctx, variant):
    """Variants 3 (fp32 FMA) and 4 (tf32 tensor-core pieces) decide most pairs from an estimate with an error bound;
    anything inside the bound must fall through to the exact fp64 test.  Trees built to sit inside that bound: clusters whose members differ by 1e-13..1e-7
    (indistinguishable in fp32), coordinates offset by 1e3 / 1e6 (huge cancellation), radius shells at r +- 1e-12."""
    rng = np.random.default_rng(77)
    ctx.set_option("nn_variant", variant)
    try:
        for trial, (offset, jitter) in enumerate(((0.0, 1e-13), (0.0, 1e-8), (1.0e3, 1e-9), (1.0e6, 1e-6), (0.0, 1e-4))):
            centres = syn.sample_configs(40, seed=500 + trial) + offset
            nodes = np.repeat(centres, 60, axis=0) + rng.normal(size=(2400, 8)) * jitter
            nodes[::60] = centres                                        # exact duplicates of a later query
            cost = np.round(rng.uniform(0, 5, len(nodes)), 2)
            queries = np.concatenate([centres[:20], centres[20:] + rng.normal(size=(20, 8)) * jitter,
                                      syn.sample_configs(24, seed=600 + trial) + offset])
            ctx.tree_reset(5)
            ctx.tree_append_batch(5, nodes, cost)
            assert np.array_equal(ctx.nn_nearest(5, queries), oracle.nearest(nodes, queries, nthreads=8)), trial
            # radius placed exactly on shells of nodes: d(node, query) computed like the reference, r = d, d +- ulp-ish
            d0 = np.sqrt(((nodes[7] - queries[0]) ** 2).sum())
            for r in (1.5, max(d0, 1e-300), np.nextafter(d0, 1.0) if d0 > 0 else 1e-3, jitter * 2.0 + 1e-15):
                got = ctx.nn_radius(5, queries, None, float(r), 20)
                want = oracle.near_topk(nodes, cost, queries, None, float(r), 20, nthreads=8)
                for a, b, name in zip(got, want, ("low", "high", "count")):
                    assert np.array_equal(a, b), (trial, r, name)
    finally:
        ctx.set_option("nn_variant", 0)


def test_tensor_core_estimate_error_is_far_inside_its_budget():
    """tools/tc_probe_f16.cu runs one 128 x 128 tile of the near-neighbour estimate (csrc/tc_umma.cuh: two tcgen05.mma over
    fp16-split operands, accumulator read back with tcgen05.ld) against fp64.  nn.cu budgets 2^-14 R^2 for the estimate's
    error because the tensor core's accumulation order is not documented; the measured error has to stay at least 8x
    below that (it is 2^-21.8 R^2), so a change of hardware or operand layout that eats into the margin is noticed."""
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "build", "tc_probe_f16")
    if not os.path.exists(exe):
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-I", os.path.join(root, "squirrel_motion_planner_b200", "csrc"),
                        "-o", exe, os.path.join(root, "tools", "tc_probe_f16.cu")], check=True, timeout=300)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "tc_probe_f16: OK" in r.stdout, r.stdout + r.stderr
    log2_err = float(re.search(r"= 2\^(-?[0-9.]+) R\^2", r.stdout).group(1))
    assert log2_err < -17.0, r.stdout


def test_tensor_core_scan_edge_cases(ctx):
    """Variant 4 (tcgen05 pre-filter on fp16 pieces, nn.cu: k_nn_tc): ragged sizes (Q not a multiple of the 256-query CTA
    tile, N not a multiple of the 128-node tile, several chunks), queries whose coordinates do not fit fp16 (zero operand
    row, every pair evaluated exactly), a NaN query, a tree that does not fit fp16 (falls back to variant 3), tiny
    coordinates (fp16 subnormal pieces) — ids and radius sets equal the oracle's in every case."""
    ctx.set_option("nn_variant", 4)
    try:
        for trial, (N, Q, scale) in enumerate(((70001, 2500, 1.0), (5003, 301, 1.0), (12345, 777, 1e-4), (9000, 513, 300.0))):
            nodes, cost = syn.random_tree(N, seed=900 + trial, cost_seed=950 + trial)
            nodes = nodes * scale
            queries = syn.sample_configs(Q, seed=970 + trial) * scale
            queries[3] = nodes[11]
            queries[5] = queries[5] + 1.0e3                # |q|^2 > 6e4: unusable as an fp16 operand
            queries[7, 2] = np.nan
            nodes[N - 1] = nodes[11]                        # tie with a later id
            ctx.tree_reset(6)
            ctx.tree_append_batch(6, nodes, cost)
            want = oracle.nearest(nodes, queries, nthreads=8)
            assert np.array_equal(ctx.nn_nearest(6, queries), want), trial
            r = 1.5 * scale
            got = ctx.nn_radius(6, queries, None, float(r), 20)
            ref = oracle.near_topk(nodes, cost, queries, None, float(r), 20, nthreads=8)
            for a, b, name in zip(got, ref, ("low", "high", "count")):
                assert np.array_equal(a, b), (trial, name)
    finally:
        ctx.set_option("nn_variant", 0)


def test_single_query_fast_path_equals_batched_path(ctx, orc, small_world):
    """Q = 1 (the planner's call shape) takes the one-launch path (k_nn_single, answers in mapped pinned memory) and
    validity calls of <= 4096 states take the zero-copy path; both must equal the batched kernels and the oracle."""
    rng = np.random.default_rng(31)
    for N in (1, 2, 33, 1025, 40000):
        nodes, cost = syn.random_tree(N, seed=700 + N, cost_seed=701 + N, xy_half=2.0)
        cost = np.round(cost, 1)
        if N > 40:
            nodes[N // 3] = nodes[5]
        ctx.tree_reset(6)
        ctx.tree_append_batch(6, nodes, cost)
        queries = syn.sample_configs(6, seed=702 + N, xy_half=2.0)
        queries[0] = nodes[min(5, N - 1)]
        want_nn = oracle.nearest(nodes, queries)
        for i, x in enumerate(queries):
            assert ctx.nn_nearest(6, x[None])[0] == want_nn[i]                       # fast path, one query at a time
            ex = int(rng.integers(-1, N))
            for r, k in ((1.5, 20), (3.0, 5), (100.0, 64), (1e-9, 20)):
                got = ctx.nn_radius(6, x[None], np.array([ex], np.int32), r, k)       # fast path
                want = oracle.near_topk(nodes, cost, x[None], np.array([ex], np.int32), r, k)
                for a, b, name in zip(got, want, ("low", "high", "count")):
                    assert np.array_equal(a, b), (N, i, r, k, name)
            got = ctx.nn_radius(6, x[None], None, 2.5, 100)                           # k > 64: batched path
            want = oracle.near_topk(nodes, cost, x[None], None, 2.5, 100)
            assert all(np.array_equal(a, b) for a, b in zip(got, want))
        # appends and updates are stream-ordered, not synchronised: the next query must still see them
        far = np.full(8, 9.0)
        nid = ctx.tree_append(6, far, 0.5)
        assert ctx.nn_nearest(6, far[None])[0] == nid
        ctx.tree_update(6, 0, far + 1e-3, 0.25)
        assert ctx.nn_nearest(6, (far + 1e-3)[None])[0] == 0
    w, origin = small_world
    ctx.set_voxel_grid(w.bits, w.dims, origin, w.res)
    orc.set_grid(w.bits, w.dims, origin, w.res)
    q = syn.state_batch(4096, seed=33, xy_half=2.2)
    for S in (1, 21, 945, 4096):
        a = unpack(ctx.check_states(q[:S], api.CHECK_MAP | api.CHECK_SELF), S)        # zero-copy path
        ctx.set_option("small_call_states", 0)
        b = unpack(ctx.check_states(q[:S], api.CHECK_MAP | api.CHECK_SELF), S)        # staged path
        ctx.set_option("small_call_states", 4096)
        assert np.array_equal(a, b)
    starts, ends = syn.edge_batch(45, seed=34, xy_half=2.2)
    m1, f1 = ctx.check_edges(starts, ends, 20, api.CHECK_MAP)
    ctx.set_option("small_call_states", 0)
    m2, f2 = ctx.check_edges(starts, ends, 20, api.CHECK_MAP)
    ctx.set_option("small_call_states", 4096)
    mo, fo = orc.check_edges(starts, ends, 20, oracle.FLAG_MAP)
    assert np.array_equal(m1, m2) and np.array_equal(f1, f2) and np.array_equal(f1, fo)


def test_radius_overflow_retry_is_exact(ctx):
    """A hit pool far too small for the batch: the host entry point halves the batch until it fits."""
    nodes, cost = syn.random_tree(4000, seed=21, cost_seed=22)
    queries = syn.sample_configs(200, seed=23)
    ctx.tree_reset(1)
    ctx.tree_append_batch(1, nodes, cost)
    ctx.set_option("nn_hit_pool_bytes", 200 * 32 * 4)                  # 32 slots per query, most need far more
    try:
        got = ctx.nn_radius(1, queries, None, 100.0, 20)
        want = oracle.near_topk(nodes, cost, queries, None, 100.0, 20, nthreads=8)
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
        assert got[2].min() == 4000
    finally:
        ctx.set_option("nn_hit_pool_bytes", 1 << 30)


def test_shim_nearest_and_near_vertices(fixture_data):
    cc = planner_shim.CollisionChecker(0)
    hp = planner_shim.BiRRTstarHotPath(cc)
    tree = planner_shim.Rrt_star_tree("START", index=3)
    hp.resetTree(tree)
    nodes, cost = syn.random_tree(600, seed=51, cost_seed=52, xy_half=1.0)
    for i in range(600):
        assert hp.insertNode(tree, planner_shim.Node(config=nodes[i], cost_reach_total=cost[i])) == i
    x = syn.sample_configs(1, seed=53, xy_half=1.0)[0]
    assert hp.find_nearest_neighbour_interpolation(tree, x).node_id == oracle.nearest(nodes, x[None])[0]
    x_new = tree.nodes[17]
    lo, hi, n = hp.find_near_vertices_interpolation(tree, x_new)
    want = oracle.near_vertices(nodes, cost, x_new.config, 17, 1.5)
    assert n == len(want) and lo == list(want[:20]) and hi == list(want[::-1][:20]) and 17 not in lo


# --------------------------------------------------------------------------------------- full-size properties
def test_full_size_config2_properties(ctx, orc):
    """2^20 states vs W256 (BASELINE.json configs[1]): entry points agree with each other, precisions agree outside
    the band, and a seeded sample of the full batch agrees with the oracle."""
    w = syn.box_world(g=256, n_boxes=64, seed=1234)
    origin = w.origin + np.array([0, 0, rd.OCTREE_Z_OFFSET])
    ctx.set_voxel_grid(w.bits, w.dims, origin, w.res)
    orc.set_grid(w.bits, w.dims, origin, w.res)
    S = 1 << 20
    E = (S + 20) // 21
    a, b = syn.edge_batch(E, seed=42)
    q = np.ascontiguousarray(syn.interpolate_edges(a, b)[:S])
    m_states = ctx.check_states(q, api.CHECK_MAP)
    m_edges, first = ctx.check_edges(a, b, 20, api.CHECK_MAP)
    assert np.array_equal(unpack(m_states, S), unpack(m_edges, S))            # idempotent across entry points
    assert np.array_equal(ctx.check_states(q, api.CHECK_MAP), m_states)       # and across runs
    bits = unpack(m_edges, E * 21).reshape(E, 21)
    assert np.array_equal(first, np.where(bits.any(1), bits.argmax(1), -1))
    sel = np.random.default_rng(0).choice(S, 40000, replace=False)
    ref, cm, _ = orc.check_states(q[sel], oracle.FLAG_MAP, clearance=True, nthreads=8)
    assert not np.any((unpack(ref, len(sel)) != unpack(m_states, S)[sel]) & (np.abs(cm) > BAND))


def test_full_size_config4_sharding_invariance(ctx, orc):
    """2^26 states vs W256 (BASELINE.json configs[3]), generated on the device (K5) so that no 4 GiB host array is
    needed: the mask of the whole batch equals the concatenation of the masks of its 8 rank shards (what the multi-GPU
    run all-gathers), both kernels agree on a 2^22 slice, and a seeded sample agrees with the oracle."""
    import torch
    w = syn.box_world(g=256, n_boxes=64, seed=1234)
    origin = w.origin + np.array([0, 0, rd.OCTREE_Z_OFFSET])
    ctx.set_voxel_grid(w.bits, w.dims, origin, w.res)
    orc.set_grid(w.bits, w.dims, origin, w.res)
    S, R = 1 << 26, 8
    stream = torch.cuda.current_stream().cuda_stream
    lo = np.array([-5.0, -5.0, -np.pi, -1.2, -1.7, -1.8, -2.4, -2.9])
    hi = np.array([5.0, 5.0, np.pi, 1.5, 2.6, 1.8, 2.4, 2.9])
    d_q = torch.empty((S, 8), dtype=torch.float64, device="cuda")
    ctx.sample_configs_dev(2026, 0, S, lo, hi, 8, d_q, None, stream)
    d_whole = torch.zeros(S // 32, dtype=torch.int32, device="cuda")
    d_parts = torch.zeros(S // 32, dtype=torch.int32, device="cuda")
    ctx.check_states_dev(d_q, S, api.CHECK_MAP, d_whole, stream)
    for r in range(R):                                                   # rank r's shard, written into its slot
        lo_s, n = r * (S // R), S // R
        ctx.check_states_dev(d_q[lo_s:], n, api.CHECK_MAP, d_parts[lo_s // 32:], stream)
    torch.cuda.synchronize()
    assert torch.equal(d_whole, d_parts)
    frac = float(torch.sum(torch.bitwise_count(d_whole).to(torch.int64)).item()) / S if hasattr(torch, "bitwise_count") else \
        float(np.unpackbits(d_whole[: 1 << 16].cpu().numpy().view(np.uint8)).mean())
    assert 0.05 < frac < 0.95
    n_slice = 1 << 22
    d_a = torch.zeros(n_slice // 32, dtype=torch.int32, device="cuda")
    ctx.set_option("k1_kernel", 1)
    ctx.check_states_dev(d_q[S // 2:], n_slice, api.CHECK_MAP, d_a, stream)
    ctx.set_option("k1_kernel", 0)
    torch.cuda.synchronize()
    assert torch.equal(d_a, d_whole[S // 64: S // 64 + n_slice // 32])
    sel = np.random.default_rng(1).choice(S, 20000, replace=False)
    sel.sort()
    q_sel = d_q[torch.from_numpy(sel).cuda()].cpu().numpy()
    got = unpack(d_whole.cpu().numpy(), S)[sel]
    ref, cm, _ = orc.check_states(q_sel, oracle.FLAG_MAP, clearance=True, nthreads=8)
    assert not np.any((unpack(ref, len(sel)) != got) & (np.abs(cm) > BAND))
    del d_q, d_whole, d_parts
    torch.cuda.empty_cache()


def test_full_size_config3_properties(ctx):
    """configs[2] at its full size — 2^20-node tree, Q = 2^16 queries (what bench.py --workload config3 times): a seeded
    256-query subset against the oracle, bit-exact, nearest and radius/top-k, both mappings."""
    N, Q = 1 << 20, 1 << 16
    nodes, cost = syn.random_tree(N, seed=7, cost_seed=9)
    queries = syn.sample_configs(Q, seed=8)
    ctx.tree_reset(0)
    ctx.tree_append_batch(0, nodes, cost)
    ids = ctx.nn_nearest(0, queries)                                          # QT mapping with the fp32 pre-filter (Q >= 2048)
    sub = np.random.default_rng(3).choice(Q, 256, replace=False)
    sub.sort()
    want = oracle.nearest(nodes, queries[sub], nthreads=8)
    assert np.array_equal(ids[sub], want)
    assert np.array_equal(ctx.nn_nearest(0, queries[sub]), want)              # WQ mapping (Q = 256)
    low, high, cnt = ctx.nn_radius(0, queries, None, 1.5, 20)
    wl, wh, wc = oracle.near_topk(nodes, cost, queries[sub], None, 1.5, 20, nthreads=8)
    assert np.array_equal(cnt[sub], wc) and np.array_equal(low[sub], wl) and np.array_equal(high[sub], wh)
    l2, h2, c2 = ctx.nn_radius(0, queries[sub], None, 1.5, 20)
    assert np.array_equal(c2, wc) and np.array_equal(l2, wl) and np.array_equal(h2, wh)
    assert (ids >= 0).all() and (ids < N).all() and cnt.max() <= N and (low[cnt > 0, 0] >= 0).all()


def test_config5_scene_w512_both_kernels(ctx, orc):
    """BASELINE.json configs[4] scene: W512 (512^3 voxels at 0.05 m, floor + 2048 boxes, seed 4321 — 16 MiB of bits, 167 MiB
    of lower-bound field, the fields no longer fit L2).  Both kernels against the oracle on >= 20 k states, map and
    map + self, and against each other bit for bit on the edge entry point."""
    w = syn.box_world(g=512, n_boxes=2048, seed=4321)
    origin = w.origin + np.array([0, 0, rd.OCTREE_Z_OFFSET])
    ctx.set_disabled_links(0)
    orc.set_disabled_links(0)
    ctx.set_voxel_grid(w.bits, w.dims, origin, w.res)
    orc.set_grid(w.bits, w.dims, origin, w.res)
    q0 = syn.state_batch(21 * 1000 + 13, seed=77, xy_half=12.0)
    try:
        # 2048 boxes leave ~1 % of uniformly drawn states free: the batch is the uniform draw plus a cloud around its free
        # states (the neighbourhood of free space is where verdicts are mixed and the bounds are least effective)
        free = q0[~unpack(ctx.check_states(q0, api.CHECK_MAP), len(q0))]
        assert 50 < len(free) < 0.05 * len(q0)
        rng = np.random.default_rng(79)
        cloud = free[rng.integers(0, len(free), 20000)] + rng.normal(0.0, 1.0, (20000, 8)) * np.array([0.12, 0.12, 0.4, 0.3, 0.3, 0.3, 0.3, 0.3])
        cloud[:, 3:] = np.clip(cloud[:, 3:], syn.ARM_LO, syn.ARM_HI)
        q = np.concatenate([q0, cloud])
        masks = {}
        for kernel in (1, 2):
            ctx.set_option("k1_kernel", kernel)
            for flags in (api.CHECK_MAP, api.CHECK_MAP | api.CHECK_SELF):
                got, ref, clr, differs = _verdict_parity(ctx, orc, q, flags)
                assert differs.sum() == 0 and 0.3 < got.mean() < 0.95
                masks[(kernel, flags)] = got
        for flags in (api.CHECK_MAP, api.CHECK_MAP | api.CHECK_SELF):
            assert np.array_equal(masks[(1, flags)], masks[(2, flags)])
        ea, eb = syn.edge_batch(4000, seed=78, xy_half=12.0)
        ctx.set_option("k1_kernel", 1)
        m1, f1 = ctx.check_edges(ea, eb, 20, api.CHECK_MAP)
        ctx.set_option("k1_kernel", 2)
        m2, f2 = ctx.check_edges(ea, eb, 20, api.CHECK_MAP)
        assert np.array_equal(m1, m2) and np.array_equal(f1, f2)
        rm, rf = orc.check_edges(ea[:300], eb[:300], 20, oracle.FLAG_MAP, nthreads=8)
        assert np.array_equal(f2[:300], rf) and np.array_equal(unpack(m2, 300 * 21), unpack(rm, 300 * 21))
    finally:
        ctx.set_option("k1_kernel", 0)
