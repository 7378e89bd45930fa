"""ctypes binding of the CPU oracle (oracle/sqmp_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this.
"Parity unpinned": see sqmp_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsqmp_oracle.so")
FLAG_SELF, FLAG_MAP = 1, 2
DOF = 8

_lib = None
DOF = 8


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "sqmp_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("sqmp_oracle.c", "sqmp_oracle.h", "Makefile"))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    assert os.path.exists(src)
    return _LIB_PATH


def build_planner() -> str:
    """oracle/sqmp_plan_oracle: the package's planner front-end (shim/sqmp_birrt.hpp) over this CPU oracle — the
    decision-parity twin of shim/sqmp_plan and the CPU arm of config 5.  Needs the package's library to be built."""
    build()
    subprocess.run(["make", "-s", "-C", _HERE, "planner"], check=True, stdout=subprocess.DEVNULL)
    return os.path.join(_HERE, "sqmp_plan_oracle")


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


# ---- oracle/_ref: the reference's OWN distance / sort code (distance_heuristics.cpp, data_structs.h) ----------
_REF_PATH = os.path.join(_HERE, "_ref", "libsqmp_ref.so")
REFERENCE_ROOT = "/root/reference"
_ref = None


def build_ref() -> str | None:
    """Compile oracle/_ref/libsqmp_ref.so from the reference sources where they lie (oracle/Makefile.ref).
    Only possible where /root/reference exists (the build container); the GPU box uses the prebuilt file."""
    if os.path.isdir(REFERENCE_ROOT):
        subprocess.run(["make", "-s", "-C", _HERE, "-f", "Makefile.ref", "REF=" + REFERENCE_ROOT], check=True,
                       stdout=subprocess.DEVNULL)
    return _REF_PATH if os.path.exists(_REF_PATH) else None


def ref_lib():
    """ctypes handle of oracle/_ref/libsqmp_ref.so, or None when it has not been built."""
    global _ref
    if _ref is not None:
        return _ref
    if not os.path.exists(_REF_PATH) and build_ref() is None:
        return None
    L = C.CDLL(_REF_PATH)
    dp, ip, i32p = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int32)
    L.ref_joint_space_distance.restype = C.c_double
    L.ref_joint_space_distance.argtypes = [dp, dp, C.c_int]
    L.ref_joint_space_distance_chain.argtypes = [ip, C.c_int, dp, dp, C.c_int, dp]
    L.ref_nearest_batch.argtypes = [dp, C.c_int64, dp, C.c_int64, i32p]
    L.ref_near_vertices.restype = C.c_int64
    L.ref_near_vertices.argtypes = [dp, dp, C.c_int64, dp, C.c_int32, C.c_double, i32p, C.c_int64]
    _ref = L
    return L


def ref_joint_space_distance(a, b):
    a, b = np.ascontiguousarray(a, np.float64), np.ascontiguousarray(b, np.float64)
    return ref_lib().ref_joint_space_distance(_p(a, C.c_double), _p(b, C.c_double), a.size)


def ref_joint_space_distance_chain(seg_types, a, b):
    t = np.ascontiguousarray(seg_types, np.int32)
    a, b = np.ascontiguousarray(a, np.float64), np.ascontiguousarray(b, np.float64)
    out = np.empty(3)
    ref_lib().ref_joint_space_distance_chain(_p(t, C.c_int), t.size, _p(a, C.c_double), _p(b, C.c_double), a.size,
                                             _p(out, C.c_double))
    return out


def ref_nearest(nodes, queries):
    nodes = np.ascontiguousarray(nodes, np.float64).reshape(-1, DOF)
    queries = np.ascontiguousarray(queries, np.float64).reshape(-1, DOF)
    ids = np.empty(queries.shape[0], dtype=np.int32)
    ref_lib().ref_nearest_batch(_p(nodes, C.c_double), nodes.shape[0], _p(queries, C.c_double), queries.shape[0],
                                _p(ids, C.c_int32))
    return ids


def ref_near_vertices(nodes, cost, query, exclude_id, r):
    nodes = np.ascontiguousarray(nodes, np.float64).reshape(-1, DOF)
    cost, query = np.ascontiguousarray(cost, np.float64), np.ascontiguousarray(query, np.float64)
    ids = np.empty(nodes.shape[0], dtype=np.int32)
    n = ref_lib().ref_near_vertices(_p(nodes, C.c_double), _p(cost, C.c_double), nodes.shape[0], _p(query, C.c_double),
                                    int(exclude_id), float(r), _p(ids, C.c_int32), ids.size)
    return ids[:n].copy()


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    dp, ip, u32p, i32p, u8p = (C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_uint32),
                               C.POINTER(C.c_int32), C.POINTER(C.c_ubyte))
    L.orc_model_create.restype = C.c_void_p
    L.orc_model_create.argtypes = [C.c_int, ip, ip, ip, dp, dp, dp, dp, ip, u8p, C.c_int, ip, dp, dp]
    L.orc_model_destroy.argtypes = [C.c_void_p]
    L.orc_set_grid.argtypes = [C.c_void_p, u32p, C.c_int, C.c_int, C.c_int, dp, C.c_double]
    L.orc_set_disabled_links.argtypes = [C.c_void_p, C.c_uint64]
    L.orc_update_transforms.argtypes = [C.c_void_p, dp, dp]
    L.orc_is_in_collision.restype = C.c_int
    L.orc_is_in_collision.argtypes = [C.c_void_p, dp, C.c_int, C.c_int]
    L.orc_state_clearance.restype = C.c_int
    L.orc_state_clearance.argtypes = [C.c_void_p, dp, C.c_uint, C.c_double, dp, dp]
    L.orc_check_states.argtypes = [C.c_void_p, dp, C.c_int64, C.c_uint, C.c_double, u32p, dp, dp, C.c_int]
    L.orc_get_collisions.argtypes = [C.c_void_p, dp, u8p, u8p]
    L.orc_interpolate.argtypes = [dp, dp, C.c_int, dp]
    L.orc_is_edge_valid.restype = C.c_int
    L.orc_is_edge_valid.argtypes = [C.c_void_p, dp, C.c_int, C.c_int, C.c_int, ip]
    L.orc_check_edges.argtypes = [C.c_void_p, dp, dp, C.c_int64, C.c_int, C.c_uint, u32p, i32p, C.c_int]
    L.orc_fk.argtypes = [C.c_void_p, C.c_int, dp, C.c_int64, dp]
    L.orc_sample_configs.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int64, C.c_int64, dp, dp, C.c_int, dp,
                                     C.POINTER(C.c_int32)]
    L.orc_sample_informed.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int64, C.c_int64, dp, dp, C.c_int, C.c_void_p, dp,
                                      C.POINTER(C.c_int32)]
    L.orc_joint_space_distance.restype = C.c_double
    L.orc_joint_space_distance.argtypes = [dp, dp, C.c_int]
    L.orc_nearest.restype = C.c_int32
    L.orc_nearest.argtypes = [dp, C.c_int64, dp]
    L.orc_nearest_batch.argtypes = [dp, C.c_int64, dp, C.c_int64, i32p, C.c_int]
    L.orc_near_vertices.restype = C.c_int64
    L.orc_near_vertices.argtypes = [dp, dp, C.c_int64, dp, C.c_int32, C.c_double, i32p, C.c_int64]
    L.orc_near_topk_batch.argtypes = [dp, dp, C.c_int64, dp, i32p, C.c_int64, C.c_double, C.c_int, i32p, i32p,
                                      i32p, C.c_int]
    L.orc_edge_cost.argtypes = [dp, C.c_int, ip, dp, dp]
    L.orc_step_towards.restype = C.c_int
    L.orc_step_towards.argtypes = [dp, dp, ip, C.c_double]
    L.orc_max_threads.restype = C.c_int
    _lib = L
    return L


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    """One reference-style checker: link table + spheres (+ grid)."""

    def __init__(self, link_table, spheres):
        L = lib()
        lt = link_table
        self.lt = lt
        self.n_links = lt.n
        sph_link = np.array([lt.index(n) for n, _, _ in spheres], dtype=np.int32)
        sph_c = _f64([c for _, c, _ in spheres]).reshape(-1, 3)
        sph_r = _f64([r for _, _, r in spheres])
        pair = np.ascontiguousarray(lt.self_pair_matrix().astype(np.uint8))
        args = [np.ascontiguousarray(lt.parent, dtype=np.int32), np.ascontiguousarray(lt.joint_type, dtype=np.int32),
                np.ascontiguousarray(lt.qidx, dtype=np.int32), _f64(lt.joint_origin), _f64(lt.joint_axis),
                _f64(lt.to_parent), _f64(lt.frame_to_tip), np.ascontiguousarray(lt.shaped, dtype=np.int32)]
        self._h = L.orc_model_create(
            lt.n, _p(args[0], C.c_int), _p(args[1], C.c_int), _p(args[2], C.c_int), _p(args[3], C.c_double),
            _p(args[4], C.c_double), _p(args[5], C.c_double), _p(args[6], C.c_double), _p(args[7], C.c_int),
            _p(pair, C.c_ubyte), len(sph_r), _p(sph_link, C.c_int), _p(sph_c, C.c_double), _p(sph_r, C.c_double))
        if not self._h:
            raise RuntimeError("orc_model_create failed")

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.orc_model_destroy(self._h)
            self._h = None

    # ---- grid / flags
    def set_grid(self, bits, dims, origin, res):
        bits = np.ascontiguousarray(bits, dtype=np.uint32)
        o = _f64(origin)
        lib().orc_set_grid(self._h, _p(bits, C.c_uint32), int(dims[0]), int(dims[1]), int(dims[2]), _p(o, C.c_double),
                           float(res))

    def set_disabled_links(self, mask: int):
        lib().orc_set_disabled_links(self._h, C.c_uint64(mask))

    # ---- state validity
    def update_transforms(self, q):
        q = _f64(q)
        out = np.empty((self.n_links, 12))
        lib().orc_update_transforms(self._h, _p(q, C.c_double), _p(out, C.c_double))
        return out

    def is_in_collision(self, q, check_self, check_map):
        q = _f64(q)
        return bool(lib().orc_is_in_collision(self._h, _p(q, C.c_double), int(check_self), int(check_map)))

    def check_states(self, q, flags=FLAG_MAP, clearance=False, cap=0.1, nthreads=1):
        q = _f64(q).reshape(-1, DOF)
        S = q.shape[0]
        mask = np.zeros((S + 31) // 32, dtype=np.uint32)
        cm = np.empty(S) if clearance else None
        cs = np.empty(S) if clearance else None
        lib().orc_check_states(self._h, _p(q, C.c_double), S, flags, cap, _p(mask, C.c_uint32),
                               _p(cm, C.c_double) if clearance else None, _p(cs, C.c_double) if clearance else None,
                               nthreads)
        return (mask, cm, cs) if clearance else mask

    def get_collisions(self, q):
        q = _f64(q)
        lm = np.zeros(self.n_links, dtype=np.uint8)
        ph = np.zeros((self.n_links, self.n_links), dtype=np.uint8)
        lib().orc_get_collisions(self._h, _p(q, C.c_double), _p(lm, C.c_ubyte), _p(ph, C.c_ubyte))
        return lm.astype(bool), ph.astype(bool)

    def is_edge_valid(self, traj, check_self, check_map):
        traj = _f64(traj).reshape(-1, DOF)
        last = C.c_int(0)
        ok = lib().orc_is_edge_valid(self._h, _p(traj, C.c_double), traj.shape[0], int(check_self), int(check_map),
                                     C.byref(last))
        return bool(ok), last.value

    def check_edges(self, q_start, q_end, num_points=20, flags=FLAG_MAP, nthreads=1):
        a, b = _f64(q_start).reshape(-1, DOF), _f64(q_end).reshape(-1, DOF)
        E = a.shape[0]
        mask = np.zeros((E * (num_points + 1) + 31) // 32, dtype=np.uint32)
        fi = np.empty(E, dtype=np.int32)
        lib().orc_check_edges(self._h, _p(a, C.c_double), _p(b, C.c_double), E, num_points, flags,
                              _p(mask, C.c_uint32), _p(fi, C.c_int32), nthreads)
        return mask, fi

    def fk(self, q, tip="hand_wrist_link"):
        q = _f64(q).reshape(-1, DOF)
        out = np.empty((q.shape[0], 7))
        lib().orc_fk(self._h, self.lt.index(tip), _p(q, C.c_double), q.shape[0], _p(out, C.c_double))
        return out


    def sample_configs(self, seed, first, n, lo, hi, max_tries=64, tip="hand_wrist_link"):
        """sampleJointConfig_JntArray (BS:3832-3878): (q[n,8], tries[n])."""
        lo, hi = _f64(lo), _f64(hi)
        q = np.empty((n, DOF))
        tries = np.empty(n, dtype=np.int32)
        lib().orc_sample_configs(self._h, self.lt.index(tip), int(seed), int(first), int(n), _p(lo, C.c_double),
                                 _p(hi, C.c_double), int(max_tries), _p(q, C.c_double), _p(tries, C.c_int32))
        return q, tries


    def sample_informed(self, seed, first, n, lo, hi, ellipse, max_tries=64, tip="hand_wrist_link"):
        """sampleJointConfigfromEllipse_JntArray (BS:3607-3829); ellipse: squirrel_motion_planner_b200.informed.InformedC."""
        lo, hi = _f64(lo), _f64(hi)
        q = np.empty((n, DOF))
        tries = np.empty(n, dtype=np.int32)
        lib().orc_sample_informed(self._h, self.lt.index(tip), C.c_uint64(int(seed)), C.c_int64(int(first)), C.c_int64(int(n)),
                                  _p(lo, C.c_double), _p(hi, C.c_double), int(max_tries), C.byref(ellipse), _p(q, C.c_double),
                                  _p(tries, C.c_int32))
        return q, tries


# ---- free functions ---------------------------------------------------------
def interpolate(start, end, num_points=20):
    a, b = _f64(start), _f64(end)
    out = np.empty((num_points + 1, DOF))
    lib().orc_interpolate(_p(a, C.c_double), _p(b, C.c_double), num_points, _p(out, C.c_double))
    return out


def joint_space_distance(a, b):
    a, b = _f64(a), _f64(b)
    return lib().orc_joint_space_distance(_p(a, C.c_double), _p(b, C.c_double), a.size)


def nearest(nodes, queries, nthreads=1):
    nodes, queries = _f64(nodes).reshape(-1, DOF), _f64(queries).reshape(-1, DOF)
    ids = np.empty(queries.shape[0], dtype=np.int32)
    lib().orc_nearest_batch(_p(nodes, C.c_double), nodes.shape[0], _p(queries, C.c_double), queries.shape[0],
                            _p(ids, C.c_int32), nthreads)
    return ids


def near_vertices(nodes, cost, query, exclude_id, r):
    nodes, cost, query = _f64(nodes).reshape(-1, DOF), _f64(cost), _f64(query)
    ids = np.empty(nodes.shape[0], dtype=np.int32)
    n = lib().orc_near_vertices(_p(nodes, C.c_double), _p(cost, C.c_double), nodes.shape[0], _p(query, C.c_double),
                                int(exclude_id), float(r), _p(ids, C.c_int32), ids.size)
    return ids[:n].copy()


def near_topk(nodes, cost, queries, exclude_id, r, k, nthreads=1):
    nodes, cost, queries = _f64(nodes).reshape(-1, DOF), _f64(cost), _f64(queries).reshape(-1, DOF)
    Q = queries.shape[0]
    ex = np.ascontiguousarray(exclude_id, dtype=np.int32) if exclude_id is not None else np.full(Q, -1, np.int32)
    low = np.empty((Q, k), dtype=np.int32)
    high = np.empty((Q, k), dtype=np.int32)
    cnt = np.empty(Q, dtype=np.int32)
    lib().orc_near_topk_batch(_p(nodes, C.c_double), _p(cost, C.c_double), nodes.shape[0], _p(queries, C.c_double),
                              _p(ex, C.c_int32), Q, float(r), int(k), _p(low, C.c_int32), _p(high, C.c_int32),
                              _p(cnt, C.c_int32), nthreads)
    return low, high, cnt


def edge_cost(traj, is_revolute, weights=None):
    traj = _f64(traj).reshape(-1, DOF)
    rev = np.ascontiguousarray(is_revolute, dtype=np.int32)
    w = _f64(np.ones(DOF) if weights is None else weights)
    out = np.empty(3)
    lib().orc_edge_cost(_p(traj, C.c_double), traj.shape[0], _p(rev, C.c_int), _p(w, C.c_double), _p(out, C.c_double))
    return out


def step_towards(nn, x_rand, is_revolute, factor):
    nn = _f64(nn)
    x = _f64(x_rand).copy()
    rev = np.ascontiguousarray(is_revolute, dtype=np.int32)
    reached = lib().orc_step_towards(_p(nn, C.c_double), _p(x, C.c_double), _p(rev, C.c_int), float(factor))
    return bool(reached), x


def max_threads():
    return lib().orc_max_threads()
