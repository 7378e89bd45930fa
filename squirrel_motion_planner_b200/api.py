"""ctypes binding of libsqmp_b200.so (include/sqmp_b200.h) — thin, no arithmetic of its own.

Python is only the harness language here (tests, bench, multi-GPU launch via torch.distributed); a C++
planner binds the same C ABI directly.  There is no CPU fallback: if the shared library
is missing or no B200 is visible, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import robot_desc as rd

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SQMP_B200_LIB") or os.path.join(_HERE, "libsqmp_b200.so")   # override: kernel A/B builds (tools/)

DOF = 8
CHECK_SELF, CHECK_MAP = 1, 2
MAX_SEG, MAX_SPHERES, MAX_LINKS = 8, 1024, 64
MAX_TREES = 1024

EXPORTS = [
    "sqmp_create", "sqmp_destroy", "sqmp_last_error", "sqmp_strerror", "sqmp_version", "sqmp_set_voxel_grid",
    "sqmp_share_voxel_grid", "sqmp_set_disabled_links", "sqmp_check_states", "sqmp_check_edges", "sqmp_get_collisions", "sqmp_fk",
    "sqmp_tree_reset", "sqmp_tree_append", "sqmp_tree_append_batch", "sqmp_tree_update", "sqmp_tree_update_costs",
    "sqmp_tree_size", "sqmp_nn_nearest", "sqmp_nn_radius", "sqmp_check_states_dev", "sqmp_check_edges_dev",
    "sqmp_fk_dev", "sqmp_sample_configs", "sqmp_sample_configs_dev", "sqmp_sample_informed", "sqmp_sample_informed_dev", "sqmp_check_states_gather_dev", "sqmp_gather_signal_dev", "sqmp_gather_wait_dev", "sqmp_nn_nearest_multi", "sqmp_nn_radius_multi", "sqmp_nn_nearest_dev", "sqmp_nn_radius_dev", "sqmp_kernel_launches", "sqmp_set_option",
]


class SqmpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libsqmp_b200 error %d: %s" % (code, msg))
        self.code = code


class RobotDescC(C.Structure):
    _fields_ = [
        ("n_seg", C.c_int32),
        ("seg_type", C.c_int32 * MAX_SEG),
        ("seg_qidx", C.c_int32 * MAX_SEG),
        ("seg_pre", (C.c_double * 12) * MAX_SEG),
        ("seg_axis", (C.c_double * 3) * MAX_SEG),
        ("root_offset", C.c_double * 3),
        ("ee_tip", C.c_double * 12),
        ("n_spheres", C.c_int32),
        ("sph_seg", C.POINTER(C.c_int32)),
        ("sph_link", C.POINTER(C.c_int32)),
        ("sph_center", C.POINTER(C.c_double)),
        ("sph_radius", C.POINTER(C.c_double)),
        ("link_pair_mask", C.c_uint64 * MAX_LINKS),
        ("self_const_collision", C.c_int32),
    ]


MAX_PEERS = 8


class PeerGatherC(C.Structure):
    """sqmp_peer_gather (include/sqmp_b200.h): peer-mapped gather buffers / flag arrays of the ranks of one node."""
    _fields_ = [
        ("n_peers", C.c_int32),
        ("rank", C.c_int32),
        ("mask", C.c_void_p * MAX_PEERS),
        ("flag", C.c_void_p * MAX_PEERS),
        ("word_offset", C.c_int64),
    ]


_lib = None


def load_library():
    """dlopen the in-tree CUDA library; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libsqmp_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C squirrel_motion_planner_b200/csrc`); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, dp, i32p, u32p, u8p = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
    i64 = C.c_int64
    L.sqmp_create.restype = vp
    L.sqmp_create.argtypes = [C.c_int, C.POINTER(RobotDescC), C.POINTER(C.c_int)]
    L.sqmp_destroy.argtypes = [vp]
    L.sqmp_last_error.restype = C.c_char_p
    L.sqmp_last_error.argtypes = [vp]
    L.sqmp_strerror.restype = C.c_char_p
    L.sqmp_strerror.argtypes = [C.c_int]
    L.sqmp_version.restype = C.c_char_p
    L.sqmp_set_voxel_grid.argtypes = [vp, u32p, C.c_int, C.c_int, C.c_int, dp, C.c_double]
    L.sqmp_set_disabled_links.argtypes = [vp, C.c_uint64]
    L.sqmp_check_states.argtypes = [vp, vp, i64, C.c_uint, vp]
    L.sqmp_check_edges.argtypes = [vp, vp, vp, i64, C.c_int, C.c_uint, vp, vp]
    L.sqmp_get_collisions.argtypes = [vp, dp, u8p, u8p]
    L.sqmp_fk.argtypes = [vp, vp, i64, vp]
    L.sqmp_tree_reset.argtypes = [vp, C.c_int]
    L.sqmp_tree_append.argtypes = [vp, C.c_int, dp, C.c_double, i32p]
    L.sqmp_tree_append_batch.argtypes = [vp, C.c_int, vp, vp, i64, i32p]
    L.sqmp_tree_update.argtypes = [vp, C.c_int, C.c_int32, dp, C.c_double]
    L.sqmp_tree_update_costs.argtypes = [vp, C.c_int, vp, vp, i64]
    L.sqmp_tree_size.argtypes = [vp, C.c_int, C.POINTER(i64)]
    L.sqmp_nn_nearest.argtypes = [vp, C.c_int, vp, i64, vp]
    L.sqmp_nn_radius.argtypes = [vp, C.c_int, vp, vp, i64, C.c_double, C.c_int, vp, vp, vp]
    L.sqmp_check_states_dev.argtypes = [vp, vp, i64, C.c_uint, vp, vp]
    L.sqmp_check_edges_dev.argtypes = [vp, vp, vp, i64, C.c_int, C.c_uint, vp, vp, vp]
    L.sqmp_fk_dev.argtypes = [vp, vp, i64, vp, vp]
    L.sqmp_check_states_gather_dev.argtypes = [vp, vp, i64, C.c_uint, C.POINTER(PeerGatherC), vp]
    L.sqmp_gather_signal_dev.argtypes = [vp, C.POINTER(PeerGatherC), C.c_uint32, vp]
    L.sqmp_gather_wait_dev.argtypes = [vp, vp, C.c_int, C.c_uint32, vp, vp]
    L.sqmp_nn_nearest_multi.argtypes = [vp, vp, vp, i64, vp]
    L.sqmp_nn_radius_multi.argtypes = [vp, vp, vp, vp, i64, C.c_double, C.c_int, vp, vp, vp]
    L.sqmp_sample_configs.argtypes = [vp, C.c_uint64, i64, i64, dp, dp, C.c_int, vp, vp]
    L.sqmp_sample_configs_dev.argtypes = [vp, C.c_uint64, i64, i64, dp, dp, C.c_int, vp, vp, vp]
    L.sqmp_sample_informed.argtypes = [vp, C.c_uint64, i64, i64, dp, dp, C.c_int, vp, vp, vp]
    L.sqmp_sample_informed_dev.argtypes = [vp, C.c_uint64, i64, i64, dp, dp, C.c_int, vp, vp, vp, vp]
    L.sqmp_nn_nearest_dev.argtypes = [vp, C.c_int, vp, i64, vp, vp]
    L.sqmp_nn_radius_dev.argtypes = [vp, C.c_int, vp, vp, i64, C.c_double, C.c_int, vp, vp, vp, vp, vp]
    L.sqmp_kernel_launches.restype = i64
    L.sqmp_kernel_launches.argtypes = [vp]
    L.sqmp_set_option.argtypes = [vp, C.c_char_p, i64]
    for fn in EXPORTS:
        if getattr(L, fn).restype is C.c_int:
            pass
    _lib = L
    return L


def _ptr(a):
    """Raw address of a numpy array or a torch tensor (host pinned or device)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


def pack_desc(cr: rd.CompiledRobot):
    """CompiledRobot -> (RobotDescC, keepalive arrays)."""
    d = RobotDescC()
    d.n_seg = cr.n_seg
    for s in range(cr.n_seg):
        d.seg_type[s] = int(cr.seg_type[s])
        d.seg_qidx[s] = int(cr.seg_qidx[s])
        for i in range(12):
            d.seg_pre[s][i] = float(cr.seg_pre[s][i])
        for i in range(3):
            d.seg_axis[s][i] = float(cr.seg_axis[s][i])
    for i in range(3):
        d.root_offset[i] = float(cr.root_offset[i])
    for i in range(12):
        d.ee_tip[i] = float(cr.ee_tip[i])
    keep = dict(
        sph_seg=np.ascontiguousarray(cr.sph_seg, np.int32), sph_link=np.ascontiguousarray(cr.sph_link, np.int32),
        sph_center=np.ascontiguousarray(cr.sph_center, np.float64).reshape(-1),
        sph_radius=np.ascontiguousarray(cr.sph_radius, np.float64))
    d.n_spheres = cr.n_spheres
    d.sph_seg = keep["sph_seg"].ctypes.data_as(C.POINTER(C.c_int32))
    d.sph_link = keep["sph_link"].ctypes.data_as(C.POINTER(C.c_int32))
    d.sph_center = keep["sph_center"].ctypes.data_as(C.POINTER(C.c_double))
    d.sph_radius = keep["sph_radius"].ctypes.data_as(C.POINTER(C.c_double))
    for i in range(MAX_LINKS):
        d.link_pair_mask[i] = int(cr.link_pair_mask[i])
    d.self_const_collision = int(cr.self_const_collision)
    return d, keep


class Context:
    """One sqmp_ctx: robot description + resident voxel grid + tree mirrors on one GPU."""

    def __init__(self, compiled: rd.CompiledRobot, device: int = 0):
        self._L = load_library()
        self.compiled = compiled
        desc, self._keep = pack_desc(compiled)
        err = C.c_int(0)
        self._h = self._L.sqmp_create(int(device), C.byref(desc), C.byref(err))
        if not self._h:
            raise SqmpError(err.value, self._L.sqmp_strerror(err.value).decode())
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._L.sqmp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise SqmpError(rc, self._L.sqmp_last_error(self._h).decode() or self._L.sqmp_strerror(rc).decode())

    # ---- configuration
    def set_option(self, name: str, value: int):
        self._ck(self._L.sqmp_set_option(self._h, name.encode(), int(value)))

    @property
    def kernel_launches(self) -> int:
        return int(self._L.sqmp_kernel_launches(self._h))

    def set_voxel_grid(self, bits, dims, origin, res):
        bits = np.ascontiguousarray(bits, dtype=np.uint32)
        o = np.ascontiguousarray(origin, dtype=np.float64)
        nvox = int(dims[0]) * int(dims[1]) * int(dims[2])
        if bits.size < (nvox + 31) // 32:
            raise ValueError("bits array too short for the grid")
        self._ck(self._L.sqmp_set_voxel_grid(self._h, bits.ctypes.data_as(C.POINTER(C.c_uint32)), int(dims[0]),
                                             int(dims[1]), int(dims[2]), o.ctypes.data_as(C.POINTER(C.c_double)),
                                             float(res)))

    def set_disabled_links(self, mask: int):
        self._ck(self._L.sqmp_set_disabled_links(self._h, C.c_uint64(mask)))

    # ---- host-buffer entry points (numpy in, numpy out; pinned torch tensors also accepted as inputs)
    def check_states(self, q, flags=CHECK_MAP, out=None):
        if isinstance(q, np.ndarray):
            q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, DOF)
            S = q.shape[0]
        else:
            S = q.numel() // DOF
        mask = out if out is not None else np.zeros((S + 31) // 32, dtype=np.uint32)
        self._ck(self._L.sqmp_check_states(self._h, _ptr(q), S, flags, _ptr(mask)))
        return mask

    def check_edges(self, q_start, q_end, num_points=20, flags=CHECK_MAP, want_first=True, out_mask=None, out_first=None):
        """q_start / q_end: numpy arrays or pinned torch tensors [E, 8]; out_mask / out_first: optional pinned torch tensors
        (int32) the results are written to — with pageable outputs every stage of the pipeline waits for its copies."""
        if isinstance(q_start, np.ndarray) or not hasattr(q_start, "data_ptr"):
            a = np.ascontiguousarray(q_start, dtype=np.float64).reshape(-1, DOF)
            b = np.ascontiguousarray(q_end, dtype=np.float64).reshape(-1, DOF)
            E = a.shape[0]
        else:
            a, b = q_start, q_end
            E = a.numel() // DOF
        mask = out_mask if out_mask is not None else np.zeros((E * (num_points + 1) + 31) // 32, dtype=np.uint32)
        first = out_first if out_first is not None else (np.empty(E, dtype=np.int32) if want_first else None)
        self._ck(self._L.sqmp_check_edges(self._h, _ptr(a), _ptr(b), E, num_points, flags, _ptr(mask), _ptr(first)))
        return mask, first

    def get_collisions(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(DOF)
        """(link_map_hit[64], link_pair_hit[64, 64]): links that hit the map; enabled link pairs (i < j) that collide."""
        lm = np.zeros(MAX_LINKS, dtype=np.uint8)
        ph = np.zeros((MAX_LINKS, MAX_LINKS), dtype=np.uint8)
        self._ck(self._L.sqmp_get_collisions(self._h, q.ctypes.data_as(C.POINTER(C.c_double)),
                                             lm.ctypes.data_as(C.POINTER(C.c_uint8)),
                                             ph.ctypes.data_as(C.POINTER(C.c_uint8))))
        return lm.astype(bool), ph.astype(bool)

    def fk(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, DOF)
        out = np.empty((q.shape[0], 7), dtype=np.float64)
        self._ck(self._L.sqmp_fk(self._h, _ptr(q), q.shape[0], _ptr(out)))
        return out

    def nn_nearest_multi(self, trees, queries):
        """Query i against tree trees[i] (sqmp_nn_nearest_multi)."""
        q = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, DOF)
        t = np.ascontiguousarray(trees, dtype=np.int32).reshape(-1)
        ids = np.empty(q.shape[0], dtype=np.int32)
        self._ck(self._L.sqmp_nn_nearest_multi(self._h, _ptr(t), _ptr(q), q.shape[0], _ptr(ids)))
        return ids

    def nn_radius_multi(self, trees, queries, exclude, r, k):
        q = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, DOF)
        t = np.ascontiguousarray(trees, dtype=np.int32).reshape(-1)
        ex = None if exclude is None else np.ascontiguousarray(exclude, dtype=np.int32).reshape(-1)
        low = np.empty((q.shape[0], k), dtype=np.int32)
        high = np.empty((q.shape[0], k), dtype=np.int32)
        cnt = np.empty(q.shape[0], dtype=np.int32)
        self._ck(self._L.sqmp_nn_radius_multi(self._h, _ptr(t), _ptr(q), _ptr(ex), q.shape[0], float(r), int(k),
                                              _ptr(low), _ptr(high), _ptr(cnt)))
        return low, high, cnt

    def share_voxel_grid(self, other: "Context"):
        """Use the voxel fields of `other` (same device, same robot) instead of building an own copy."""
        self._L.sqmp_share_voxel_grid.argtypes = [C.c_void_p, C.c_void_p]
        self._ck(self._L.sqmp_share_voxel_grid(self._h, other._h))

    def sample_configs(self, seed, first, n, lo, hi, max_tries=64):
        """sampleJointConfig_JntArray (BS:3832-3878) batched on the device: (q[n,8], tries[n])."""
        lo = np.ascontiguousarray(lo, dtype=np.float64).reshape(DOF)
        hi = np.ascontiguousarray(hi, dtype=np.float64).reshape(DOF)
        q = np.empty((n, DOF), dtype=np.float64)
        tries = np.empty(n, dtype=np.int32)
        dp = C.POINTER(C.c_double)
        self._ck(self._L.sqmp_sample_configs(self._h, int(seed), int(first), int(n), lo.ctypes.data_as(dp),
                                             hi.ctypes.data_as(dp), int(max_tries), _ptr(q), _ptr(tries)))
        return q, tries

    def sample_informed(self, seed, first, n, lo, hi, ellipse, max_tries=64):
        """sampleJointConfigfromEllipse_JntArray (BS:3607-3829) batched on the device; ellipse: informed.InformedC."""
        lo = np.ascontiguousarray(lo, dtype=np.float64).reshape(DOF)
        hi = np.ascontiguousarray(hi, dtype=np.float64).reshape(DOF)
        q = np.empty((n, DOF), dtype=np.float64)
        tries = np.empty(n, dtype=np.int32)
        dp = C.POINTER(C.c_double)
        self._ck(self._L.sqmp_sample_informed(self._h, C.c_uint64(int(seed)), C.c_int64(int(first)), C.c_int64(int(n)),
                                              lo.ctypes.data_as(dp), hi.ctypes.data_as(dp), int(max_tries), C.byref(ellipse),
                                              _ptr(q), _ptr(tries)))
        return q, tries

    def sample_configs_dev(self, seed, first, n, lo, hi, max_tries, d_q, d_tries, stream=0):
        lo = np.ascontiguousarray(lo, dtype=np.float64).reshape(DOF)
        hi = np.ascontiguousarray(hi, dtype=np.float64).reshape(DOF)
        dp = C.POINTER(C.c_double)
        self._ck(self._L.sqmp_sample_configs_dev(self._h, int(seed), int(first), int(n), lo.ctypes.data_as(dp),
                                                 hi.ctypes.data_as(dp), int(max_tries), _ptr(d_q), _ptr(d_tries),
                                                 C.c_void_p(stream)))

    # ---- tree mirror
    def tree_reset(self, tree):
        self._ck(self._L.sqmp_tree_reset(self._h, tree))

    def tree_append(self, tree, q, cost):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(DOF)
        out = C.c_int32(-1)
        self._ck(self._L.sqmp_tree_append(self._h, tree, q.ctypes.data_as(C.POINTER(C.c_double)), float(cost),
                                          C.byref(out)))
        return out.value

    def tree_append_batch(self, tree, q, cost):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, DOF)
        cost = np.ascontiguousarray(cost, dtype=np.float64).reshape(-1)
        assert cost.size == q.shape[0]
        out = C.c_int32(-1)
        self._ck(self._L.sqmp_tree_append_batch(self._h, tree, _ptr(q), _ptr(cost), q.shape[0], C.byref(out)))
        return out.value

    def tree_update(self, tree, node_id, q, cost):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(DOF)
        self._ck(self._L.sqmp_tree_update(self._h, tree, int(node_id), q.ctypes.data_as(C.POINTER(C.c_double)),
                                          float(cost)))

    def tree_update_costs(self, tree, ids, cost):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        cost = np.ascontiguousarray(cost, dtype=np.float64)
        self._ck(self._L.sqmp_tree_update_costs(self._h, tree, _ptr(ids), _ptr(cost), ids.size))

    def tree_size(self, tree):
        n = C.c_int64(0)
        self._ck(self._L.sqmp_tree_size(self._h, tree, C.byref(n)))
        return n.value

    # ---- near neighbours
    def nn_nearest(self, tree, queries):
        qs = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, DOF)
        ids = np.empty(qs.shape[0], dtype=np.int32)
        self._ck(self._L.sqmp_nn_nearest(self._h, tree, _ptr(qs), qs.shape[0], _ptr(ids)))
        return ids

    def nn_radius(self, tree, queries, exclude_id, r, k):
        qs = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, DOF)
        Q = qs.shape[0]
        ex = None if exclude_id is None else np.ascontiguousarray(exclude_id, dtype=np.int32)
        low = np.empty((Q, k), dtype=np.int32)
        high = np.empty((Q, k), dtype=np.int32)
        cnt = np.empty(Q, dtype=np.int32)
        self._ck(self._L.sqmp_nn_radius(self._h, tree, _ptr(qs), _ptr(ex), Q, float(r), int(k), _ptr(low), _ptr(high),
                                        _ptr(cnt)))
        return low, high, cnt

    # ---- device-pointer entry points (torch CUDA tensors; stream = torch.cuda.current_stream().cuda_stream)
    def check_states_dev(self, d_q, S, flags, d_mask, stream=0):
        self._ck(self._L.sqmp_check_states_dev(self._h, _ptr(d_q), int(S), flags, _ptr(d_mask), C.c_void_p(stream)))

    def check_edges_dev(self, d_a, d_b, E, num_points, flags, d_mask, d_first=None, stream=0):
        self._ck(self._L.sqmp_check_edges_dev(self._h, _ptr(d_a), _ptr(d_b), int(E), int(num_points), flags,
                                              _ptr(d_mask), _ptr(d_first), C.c_void_p(stream)))

    def check_states_gather_dev(self, d_q, S, flags, peers: "PeerGatherC", stream=0):
        """Validity kernel with the mask all-gather fused in: words go straight to every rank's gather buffer."""
        self._ck(self._L.sqmp_check_states_gather_dev(self._h, _ptr(d_q), int(S), flags, C.byref(peers), C.c_void_p(stream)))

    def gather_signal_dev(self, peers: "PeerGatherC", step, stream=0):
        self._ck(self._L.sqmp_gather_signal_dev(self._h, C.byref(peers), int(step), C.c_void_p(stream)))

    def gather_wait_dev(self, d_flags_ptr, n_peers, step, d_status, stream=0):
        self._ck(self._L.sqmp_gather_wait_dev(self._h, C.c_void_p(int(d_flags_ptr)), int(n_peers), int(step), _ptr(d_status),
                                              C.c_void_p(stream)))

    def fk_dev(self, d_q, S, d_out, stream=0):
        self._ck(self._L.sqmp_fk_dev(self._h, _ptr(d_q), int(S), _ptr(d_out), C.c_void_p(stream)))

    def nn_nearest_dev(self, tree, d_queries, Q, d_ids, stream=0):
        self._ck(self._L.sqmp_nn_nearest_dev(self._h, tree, _ptr(d_queries), int(Q), _ptr(d_ids), C.c_void_p(stream)))

    def nn_radius_dev(self, tree, d_queries, d_exclude, Q, r, k, d_low, d_high, d_count, d_overflow, stream=0):
        self._ck(self._L.sqmp_nn_radius_dev(self._h, tree, _ptr(d_queries), _ptr(d_exclude), int(Q), float(r), int(k),
                                            _ptr(d_low), _ptr(d_high), _ptr(d_count), _ptr(d_overflow),
                                            C.c_void_p(stream)))


def default_context(device: int = 0) -> Context:
    """Context for the SQUIRREL Robotino from the committed fixture (URDF-derived constants + sphere model)."""
    lt, fx = rd.load_fixture()
    spheres = [(n, tuple(c), r) for n, c, r in fx["spheres"]]
    return Context(rd.compile_robot(lt, spheres), device)
