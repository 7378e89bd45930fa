"""Informed (ellipse) sampling set-up — host side of BiRRTstarPlanner::jointConfigEllipseInitialization
(birrt_star.cpp:3472-3604) and of the diagonal "L" of sampleJointConfigfromEllipse_JntArray (birrt_star.cpp:3709-3752).

Per joint class (revolute / prismatic) the reference builds
  centre = (start + goal) / 2,   a1 = (goal - start) / c_min   (c_min = the class's share of the start node's heuristic),
  M = a1 e1^T,  U S V^T = JacobiSVD(M),  C = U diag(1, .., 1, det(U) det(V^T)) V                      (BS:3535-3568)
i.e. a ROTATION whose first column is a1.  Eigen is not available here and the other columns of C are not determined by
M at all (its null space is (n-1)-dimensional: they come out of Eigen's sweep order), so C is built as the Householder
reflection that maps e1 to a1 with its last column negated — orthogonal, det +1, first column a1: every property the
reference's C has.  For the two prismatic joints that rotation is unique, so C_pri equals the reference's bit for bit
up to rounding.  `ctypes_struct` packs the result as sqmp_informed (include/sqmp_b200.h).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

DOF = 8


class InformedC(C.Structure):
    _fields_ = [
        ("n_rev", C.c_int32), ("n_pri", C.c_int32),
        ("rev_idx", C.c_int32 * DOF), ("pri_idx", C.c_int32 * DOF),
        ("c_rev", C.c_double * (DOF * DOF)), ("c_pri", C.c_double * (DOF * DOF)),
        ("center_rev", C.c_double * DOF), ("center_pri", C.c_double * DOF),
        ("l_rev", C.c_double * 2), ("l_pri", C.c_double * 2),
        ("env_x", C.c_double * 2), ("env_y", C.c_double * 2),
    ]


def rotation_to(a1):
    """Rotation (orthogonal, det +1) whose first column is the unit vector a1: Householder reflection e1 -> a1 with the
    last column negated (n >= 2).  a1 == e1: the identity."""
    a1 = np.asarray(a1, dtype=np.float64)
    n = len(a1)
    v = a1.copy()
    v[0] -= 1.0
    vv = 0.0
    for x in v:                          # sequential sum (the C++ twin, shim/sqmp_informed.hpp, adds in the same order)
        vv += float(x) * float(x)
    if vv == 0.0:
        return np.eye(n)
    H = np.eye(n) - (2.0 / vv) * np.outer(v, v)
    H[:, n - 1] = -H[:, n - 1]
    return H


def ellipse_init(start, goal, is_revolute):
    """jointConfigEllipseInitialization: per class index list, centre, c_min (heuristic of the start node, DH:160-217)
    and rotation C."""
    start, goal = np.asarray(start, np.float64), np.asarray(goal, np.float64)
    out = {}
    for name, sel in (("rev", [j for j in range(DOF) if is_revolute[j]]), ("pri", [j for j in range(DOF) if not is_revolute[j]])):
        idx = np.array(sel, dtype=np.int32)
        d = goal[idx] - start[idx]
        s = 0.0
        for v in d:                      # sequential sum of squares, index order (DH:175-200)
            s += v * v
        c_min = float(np.sqrt(s))
        centre = (start[idx] + goal[idx]) / 2.0
        a1 = d / c_min if c_min > 0.0 else np.eye(len(idx))[0]
        out[name] = {"idx": idx, "center": centre, "c_min": c_min, "C": rotation_to(a1) if len(idx) > 1 else np.eye(len(idx))}
    return out


def ellipse_L(c_best, c_min):
    """BS:3709-3752: L = diag(c_best / 2, sqrt(c_best^2 - c_min^2) / 2, ...)."""
    rest = c_best * c_best - c_min * c_min
    return c_best / 2.0, float(np.sqrt(rest if rest > 0.0 else 0.0)) / 2.0


def ctypes_struct(el, c_best_rev, c_best_pri, env_x=(0.0, 0.0), env_y=(0.0, 0.0)) -> InformedC:
    s = InformedC()
    for name, cb in (("rev", c_best_rev), ("pri", c_best_pri)):
        e = el[name]
        n = len(e["idx"])
        setattr(s, "n_" + name, n)
        for a in range(n):
            getattr(s, name + "_idx")[a] = int(e["idx"][a])
            getattr(s, "center_" + name)[a] = float(e["center"][a])
            for b in range(n):
                getattr(s, "c_" + name)[a * n + b] = float(e["C"][a, b])
        l0, l1 = ellipse_L(cb, e["c_min"])
        getattr(s, "l_" + name)[0] = l0
        getattr(s, "l_" + name)[1] = l1
    s.env_x[0], s.env_x[1] = env_x
    s.env_y[0], s.env_y[1] = env_y
    return s
