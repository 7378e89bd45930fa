"""Sphere-model authoring tool: build and grade CONSERVATIVE collision-sphere sets for the shapes the
reference collides (CollisionChecker's link shapes, collision_checker.hpp:263-351).

Six of the 27 shaped links carry URDF primitives (front shell box URDF:323-330, neck box 355-360, neck-pan cylinder
385-390, head cylinder 430-435, hokuyo cylinder 546-551, kinect box 1369-1374); the other 21 are STL meshes that are
not in the reference repository, so for them `sphere_model.py` declares stand-in primitives (capsules along the joint
offsets URDF:587-759, 843-846, cylinders / boxes for the body parts).  This module

  * covers a primitive completely with spheres whose union reaches at most `p` outside it (`cover_shape`): greedy set
    cover over candidate centres on a lattice inside the shape, every candidate with the largest radius the allowed
    protrusion permits (depth below the nearest face + p), checked against a dense point lattice with the lattice's
    half-diagonal as slack — so "every lattice point covered" proves that the whole volume is covered;
  * grades any sphere set against a primitive (`grade`): covered volume fraction (missed volume = possible false
    "free") and protrusion = how far the union of the spheres reaches outside the shape (possible false "collision");
  * proves a cover (`verify_cover`) with the same lattice argument the fit uses.

A shape is `(kind, dims, origin)`: `("box", (x, y, z), o)`, `("cylinder", (r, l), o)` (axis z), `("capsule", (a, b, R), o)`
(segment a-b, radius R), all in the link's collision frame, centred at / offset by `o`.
Host-side numpy/scipy only; nothing here runs on the planner's hot path.
"""
from __future__ import annotations

import heapq
import math
import re

import numpy as np


# ---------------------------------------------------------------------------------------------- shapes
def parse_shape(desc: str):
    """'box x y z' -> ('box', (x, y, z), (0,0,0));  'cylinder r=.. l=..' -> ('cylinder', (r, l), (0,0,0));  else None."""
    if desc is None:
        return None
    if desc.startswith("box"):
        return "box", tuple(float(v) for v in desc.split()[1:4]), (0.0, 0.0, 0.0)
    m = re.match(r"cylinder r=([0-9.eE+-]+) l=([0-9.eE+-]+)", desc)
    if m:
        return "cylinder", (float(m.group(1)), float(m.group(2))), (0.0, 0.0, 0.0)
    return None


def _norm_shape(shape):
    kind, dims = shape[0], shape[1]
    origin = np.asarray(shape[2] if len(shape) > 2 else (0.0, 0.0, 0.0), dtype=np.float64)
    return kind, dims, origin


def inside_depth(shape, p):
    """Signed depth of points p [n, 3] below the surface of the shape: > 0 inside (distance to the nearest face),
    < 0 outside (minus the distance to the shape)."""
    kind, dims, origin = _norm_shape(shape)
    q = np.asarray(p, dtype=np.float64).reshape(-1, 3) - origin
    if kind == "box":
        h = np.array(dims) / 2.0
        d = np.abs(q) - h
        outside = np.linalg.norm(np.maximum(d, 0.0), axis=1)
        return np.where(outside > 0.0, -outside, -d.max(axis=1))
    if kind == "cylinder":
        r, l = dims
        dr = np.hypot(q[:, 0], q[:, 1]) - r
        dz = np.abs(q[:, 2]) - l / 2.0
        outside = np.hypot(np.maximum(dr, 0.0), np.maximum(dz, 0.0))
        return np.where(outside > 0.0, -outside, -np.maximum(dr, dz))
    if kind == "capsule":
        a, b, R = dims
        return R - segment_distance(q, a, b)
    if kind == "union":   # parts given in the same frame; exact outside, a lower bound of the depth inside
        return np.max([inside_depth(part, q) for part in dims], axis=0)
    raise ValueError(kind)


def bounding_box(shape):
    kind, dims, origin = _norm_shape(shape)
    if kind == "union":
        boxes = [bounding_box(part) for part in dims]
        return origin + np.min([b[0] for b in boxes], axis=0), origin + np.max([b[1] for b in boxes], axis=0)
    if kind == "box":
        h = np.array(dims) / 2.0
        return origin - h, origin + h
    if kind == "cylinder":
        h = np.array([dims[0], dims[0], dims[1] / 2.0])
        return origin - h, origin + h
    a, b, R = dims
    a, b = np.asarray(a, float), np.asarray(b, float)
    return origin + np.minimum(a, b) - R, origin + np.maximum(a, b) + R


def segment_distance(p, a, b):
    """Distance of points p [n, 3] from the segment a-b (the joint offset that bounds a mesh link)."""
    p, a, b = np.asarray(p, float).reshape(-1, 3), np.asarray(a, float), np.asarray(b, float)
    ab = b - a
    t = np.clip(((p - a) @ ab) / max(float(ab @ ab), 1e-30), 0.0, 1.0)
    return np.linalg.norm(p - (a + t[:, None] * ab), axis=1)


def _lattice_points(shape, h):
    """Points of a lattice of pitch <= h over the bounding box that lie in the shape or within half a lattice diagonal
    of it, pulled onto the shape where they lie outside (box / cylinder: projection; capsule: kept, they only make the
    check stricter).  Returns (points, slack) with slack = half the lattice diagonal."""
    lo, hi = bounding_box(shape)
    ax = [np.linspace(lo[a], hi[a], max(int(math.ceil((hi[a] - lo[a]) / h)), 1) + 1) for a in range(3)]
    pitch = np.array([ax[a][1] - ax[a][0] for a in range(3)])
    slack = float(np.linalg.norm(pitch)) / 2.0
    P = np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
    d = inside_depth(shape, P)
    P = P[d >= -slack]
    kind, dims, origin = _norm_shape(shape)
    if kind == "cylinder":   # rim points -> onto the disc (the projection of a lattice cell's corner is in the shape)
        q = P - origin
        rad = np.hypot(q[:, 0], q[:, 1])
        sc = np.where(rad > dims[0], dims[0] / np.maximum(rad, 1e-30), 1.0)
        q[:, 0] *= sc
        q[:, 1] *= sc
        P = q + origin
    return P, slack


# ---------------------------------------------------------------------------------------------- fitting
def cover_capsule(shape, p):
    """Sphere chain of radius R + p along the segment, spaced so that consecutive spheres meet on the capsule's surface
    (2 sqrt((R+p)^2 - R^2) apart, end spheres on the end points): covers the capsule, protrudes exactly p."""
    kind, (a, b, R), origin = _norm_shape(shape)
    a, b = np.asarray(a, float) + origin, np.asarray(b, float) + origin
    r = R + p
    L = float(np.linalg.norm(b - a))
    s = 2.0 * math.sqrt(max(r * r - R * R, 1e-12))
    n = max(int(math.ceil(L / s - 1e-9)), 0)
    t = np.linspace(0.0, 1.0, n + 1) if n > 0 else np.array([0.0] if L == 0.0 else [0.0, 1.0])
    c = a[None, :] + t[:, None] * (b - a)[None, :]
    return c, np.full(len(c), r)


def cover_shape(shape, p, p_bottom=None, pitch=0.0125, h=0.0065, r_cap=0.25, r_min=0.02):
    """Spheres (centres [n, 3], radii [n]) whose union contains the whole primitive and reaches at most `p` outside it
    (`p_bottom`: tighter limit below the -z face — the floor slab is 0.04 m under the robot's lowest shape).

    Candidates: centres on a lattice of `pitch` inside the shape, each with radius min(depth + p, r_cap) (so its own
    protrusion is exactly <= p).  Points: a lattice of pitch <= h over the shape; a candidate covers a point when the
    point lies at least half a lattice diagonal inside it, so covering every lattice point covers every point of the
    volume.  Greedy: always take the candidate that covers the most still-uncovered points (lazy priority queue)."""
    from scipy.spatial import cKDTree

    kind, dims, origin = _norm_shape(shape)
    if kind == "capsule":
        return cover_capsule(shape, p)
    if kind == "union":
        parts = [cover_shape(part, p, p_bottom, pitch, h, r_cap, r_min) for part in dims]
        return np.concatenate([c for c, _ in parts]) + origin, np.concatenate([r for _, r in parts])
    lo, hi = bounding_box(shape)
    if float(np.max(hi - lo)) < 4.0 * pitch:   # tiny shape (the 1 mm dummy box URDF:355-360): one sphere around it
        c = (lo + hi) / 2.0
        return c[None, :], np.array([float(np.linalg.norm(hi - lo)) / 2.0 + min(p, 0.001)])
    P, slack = _lattice_points(shape, h)
    lo, hi = bounding_box(shape)
    ax = [np.linspace(lo[a], hi[a], max(int(round((hi[a] - lo[a]) / pitch)), 1) + 1) for a in range(3)]
    C = np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
    depth = inside_depth(shape, C)
    C, depth = C[depth >= 0.0], depth[depth >= 0.0]
    R = np.minimum(depth + p, r_cap)
    if p_bottom is not None:   # depth below the -z face limits the radius more tightly
        R = np.minimum(R, (C[:, 2] - lo[2]) + p_bottom)
    keep = R - slack > r_min * 0.5
    C, R = C[keep], R[keep]
    tree = cKDTree(P)
    cov = [np.asarray(tree.query_ball_point(C[i], R[i] - slack - 1e-5), dtype=np.int64) for i in range(len(C))]   # 1e-5: the model is stored rounded to 1e-6
    uncovered = np.ones(len(P), dtype=bool)
    heap = [(-len(c), i) for i, c in enumerate(cov)]
    heapq.heapify(heap)
    chosen = []
    while uncovered.any() and heap:
        _, i = heapq.heappop(heap)
        cnt = int(uncovered[cov[i]].sum())
        if cnt == 0:
            continue
        if heap and cnt < -heap[0][0]:
            heapq.heappush(heap, (-cnt, i))
            continue
        chosen.append(i)
        uncovered[cov[i]] = False
    if uncovered.any():
        raise ValueError("cover_shape: %d lattice points cannot be covered with protrusion %.3f (pitch %.4f)" %
                         (int(uncovered.sum()), p, pitch))
    # drop spheres made redundant by later choices (reverse pass)
    count = np.zeros(len(P), dtype=np.int32)
    for i in chosen:
        count[cov[i]] += 1
    kept = []
    for i in sorted(chosen, key=lambda i: len(cov[i])):
        if np.all(count[cov[i]] >= 2):
            count[cov[i]] -= 1
        else:
            kept.append(i)
    kept.sort(key=lambda i: (-R[i], C[i][2], C[i][1], C[i][0]))
    return C[kept], R[kept]


def verify_cover(shape, centres, radii, h=0.0065):
    """Number of lattice points (pitch <= h) of the shape NOT at least half a lattice diagonal inside some sphere.
    0 proves that the union of the spheres contains the whole shape."""
    from scipy.spatial import cKDTree

    centres = np.asarray(centres, dtype=np.float64).reshape(-1, 3)
    radii = np.asarray(radii, dtype=np.float64).reshape(-1)
    kind, dims, origin = _norm_shape(shape)
    if kind == "union":
        return sum(verify_cover(_shift(part, origin), centres, radii, h) for part in dims)
    if kind == "capsule":
        # the chain construction meets the capsule's surface exactly (no slack to give away): check its invariant
        # instead — consecutive centres on the axis, end spheres on the end points, waist radius >= R
        return _capsule_chain_gaps(shape, centres, radii)
    lo, hi = bounding_box(shape)
    if float(np.max(hi - lo)) < 0.02:
        h = float(np.max(hi - lo)) / 4.0
    P, slack = _lattice_points(shape, h)
    tree = cKDTree(P)
    ok = np.zeros(len(P), dtype=bool)
    for c, r in zip(centres, radii):
        if r - slack > 0.0:
            ok[tree.query_ball_point(c, r - slack)] = True
    return int((~ok).sum())


def _shift(shape, origin):
    kind, dims, o = _norm_shape(shape)
    return kind, dims, tuple(o + np.asarray(origin, float))


def _capsule_chain_gaps(shape, centres, radii):
    """Number of violated conditions of a sphere chain that covers capsule (a, b, R): some sphere on each end point with
    radius >= R, and between consecutive on-axis spheres (sorted along the axis) the lens they share has radius >= R."""
    kind, (a, b, R), origin = _norm_shape(shape)
    a, b = np.asarray(a, float) + origin, np.asarray(b, float) + origin
    L = float(np.linalg.norm(b - a))
    u = (b - a) / L if L > 0.0 else np.array([1.0, 0.0, 0.0])
    on = [(float((c - a) @ u), float(r)) for c, r in zip(centres, radii)
          if np.linalg.norm((c - a) - ((c - a) @ u) * u) < 1e-9]
    on.sort()
    bad = 0
    if not on or abs(on[0][0]) > 1e-9 or on[0][1] < R or abs(on[-1][0] - L) > 1e-9 or on[-1][1] < R:
        bad += 1
    for (t0, r0), (t1, r1) in zip(on[:-1], on[1:]):
        d = t1 - t0
        if d <= 1e-12:
            continue
        x = (d * d + r0 * r0 - r1 * r1) / (2.0 * d)          # plane of the lens between the two spheres
        w2 = r0 * r0 - x * x
        if w2 < R * R * (1.0 - 1e-9):
            bad += 1
    return bad


# ---------------------------------------------------------------------------------------------- grading
def _sample_volume(shape, n, rng):
    kind, dims, origin = _norm_shape(shape)
    if kind == "box":
        return rng.uniform(-0.5, 0.5, size=(n, 3)) * np.array(dims) + origin
    if kind == "cylinder":
        r, l = dims
        rad = r * np.sqrt(rng.uniform(0, 1, n))
        ang = rng.uniform(0, 2 * np.pi, n)
        return np.stack([rad * np.cos(ang), rad * np.sin(ang), rng.uniform(-l / 2, l / 2, n)], axis=1) + origin
    if kind == "union":
        return np.concatenate([_sample_volume(_shift(part, origin), max(n // len(dims), 1), rng) for part in dims])
    lo, hi = bounding_box(shape)
    P = np.zeros((0, 3))
    while len(P) < n:
        Q = rng.uniform(lo, hi, size=(8 * n, 3))
        P = np.concatenate([P, Q[inside_depth(shape, Q) >= 0.0]])
    return P[:n]


def _extreme_points(shape):
    """Corners / rim points: where a cover is most likely to miss."""
    kind, dims, origin = _norm_shape(shape)
    if kind == "box":
        h = np.array(dims) / 2.0
        return np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]) * h + origin
    if kind == "cylinder":
        ang = np.linspace(0, 2 * np.pi, 32, endpoint=False)
        return np.concatenate([np.stack([dims[0] * np.cos(ang), dims[0] * np.sin(ang), np.full(32, z)], axis=1)
                               for z in (-dims[1] / 2, dims[1] / 2)]) + origin
    if kind == "union":
        return np.concatenate([_extreme_points(_shift(part, origin)) for part in dims])
    a, b, R = dims
    a, b = np.asarray(a, float), np.asarray(b, float)
    L = float(np.linalg.norm(b - a))
    u = (b - a) / L if L > 0.0 else np.array([1.0, 0.0, 0.0])
    return np.array([a - R * u, b + R * u]) + origin


def protrusion(shape, centres, radii):
    """How far the union of the spheres reaches outside the shape: max over spheres of (radius - depth of the centre
    below the surface) — the farthest point of a sphere whose centre lies at depth d inside a convex shape is r - d
    outside (along the normal of the nearest face); for a centre outside it is distance + r."""
    centres = np.asarray(centres, dtype=np.float64).reshape(-1, 3)
    radii = np.asarray(radii, dtype=np.float64).reshape(-1)
    return float(np.maximum(radii - inside_depth(shape, centres), 0.0).max()) if len(radii) else 0.0


def grade(shape, centres, radii, n=20000, seed=0):
    """covered: fraction of the primitive's volume inside the union of the spheres (volume sample + corners / rim
    points).  protrusion: see `protrusion`."""
    centres = np.asarray(centres, dtype=np.float64).reshape(-1, 3)
    radii = np.asarray(radii, dtype=np.float64).reshape(-1)
    rng = np.random.default_rng(seed)
    p = np.concatenate([_sample_volume(shape, n, rng), _extreme_points(shape)])
    inside = np.zeros(len(p), dtype=bool)
    for c, r in zip(centres, radii):
        inside |= np.linalg.norm(p - c, axis=1) <= r + 1e-12
    return {"covered": float(inside.mean()), "protrusion": protrusion(shape, centres, radii), "n_spheres": int(len(radii))}


def report(link_shapes, spheres, n=20000):
    """Per shaped link: its (URDF or stand-in) shape, the spheres it owns and their grade.
    link_shapes: {link: (shape, source)} from sphere_model.link_shapes();  spheres: [(link, centre, radius)]."""
    rows = []
    for name, (shape, source) in link_shapes.items():
        mine = [(c, r) for n_, c, r in spheres if n_ == name]
        row = {"link": name, "shape": shape, "source": source, "n_spheres": len(mine),
               "radii": sorted({round(r, 4) for _, r in mine})}
        if mine:
            c = np.array([c for c, _ in mine])
            r = np.array([r for _, r in mine])
            row["grade"] = grade(shape, c, r, n=n)
        rows.append(row)
    return rows
