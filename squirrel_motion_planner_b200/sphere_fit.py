"""Sphere-model authoring aid: fit and grade collision spheres against the shapes the reference
collides (CollisionChecker's link shapes, collision_checker.hpp:263-351).

Six of the 27 shaped links carry URDF primitives (front shell box URDF:323-330, neck box 355-360, neck-pan cylinder
385-390, head cylinder 430-435, hokuyo cylinder 546-551, kinect box 1369-1374); the other 21 are STL meshes that are
not in the reference repository, so for them only the joint offsets that bound each link are known (URDF:587-759,
843-846).  This module
  * fits a sphere set of a given radius to a primitive so that the primitive is completely covered (a tiling of the
    bounding box whose cells fit inside a sphere of that radius) — `fit_primitive`;
  * grades any sphere set against a primitive: covered volume fraction (missed volume = possible false "free") and
    protrusion (how far the spheres reach outside the shape = possible false "collision") — `grade`;
  * grades mesh-link spheres against the segment between the link's joint origin and its child's joint origin
    — `segment_distance`.
Host-side numpy only; nothing here runs on the planner's hot path.
"""
from __future__ import annotations

import math
import re

import numpy as np


def parse_shape(desc: str):
    """'box x y z' -> ('box', (x, y, z));  'cylinder r=.. l=..' -> ('cylinder', (r, l)) (axis z, centred);  else None."""
    if desc is None:
        return None
    if desc.startswith("box"):
        return "box", tuple(float(v) for v in desc.split()[1:4])
    m = re.match(r"cylinder r=([0-9.eE+-]+) l=([0-9.eE+-]+)", desc)
    if m:
        return "cylinder", (float(m.group(1)), float(m.group(2)))
    return None


def fit_primitive(shape, r: float):
    """Centres [n, 3] of radius-r spheres whose union contains the whole primitive (box, or z-axis cylinder, centred at the
    origin of the collision frame).  The bounding box is tiled by n_x * n_y * n_z equal cells with a sphere at every cell
    centre: a point of a cell is at most (h_x/n_x, h_y/n_y, h_z/n_z) from its centre, so the cover is complete iff that
    vector is no longer than r; the counts grow greedily on the axis with the largest cell until it is.  Cylinder: cells
    that do not touch the disc are dropped and rim centres are pulled onto the disc.  All centres lie inside the shape,
    so the protrusion is at most r."""
    kind, dims = shape
    half = np.array(dims) / 2.0 if kind == "box" else np.array([dims[0], dims[0], dims[1] / 2.0])
    if r <= 0.0:
        raise ValueError("radius must be positive")
    n = np.ones(3, dtype=int)
    while float(np.sum((half / n) ** 2)) > r * r * (1.0 - 1e-12):
        n[int(np.argmax(half / n))] += 1
    axes = [-half[a] + (2.0 * half[a] / n[a]) * (np.arange(n[a]) + 0.5) for a in range(3)]
    gx, gy, gz = np.meshgrid(*axes, indexing="ij")
    c = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], axis=1)
    if kind == "cylinder":
        cell = half / n
        near = np.maximum(np.abs(c[:, :2]) - cell[:2], 0.0)
        c = c[np.hypot(near[:, 0], near[:, 1]) <= dims[0]]
        # a kept centre outside the disc is projected onto it: the projection onto a convex set is at least as close
        # to every point of the set, so the cover stays complete and no centre lies outside the shape
        rad = np.hypot(c[:, 0], c[:, 1])
        scale = np.where(rad > dims[0], dims[0] / np.maximum(rad, 1e-30), 1.0)
        c[:, 0] *= scale
        c[:, 1] *= scale
    return c


def _sample_volume(shape, n, rng):
    kind, dims = shape
    if kind == "box":
        return rng.uniform(-0.5, 0.5, size=(n, 3)) * np.array(dims)
    r, l = dims
    rad = r * np.sqrt(rng.uniform(0, 1, n))
    ang = rng.uniform(0, 2 * np.pi, n)
    return np.stack([rad * np.cos(ang), rad * np.sin(ang), rng.uniform(-l / 2, l / 2, n)], axis=1)


def _outside_distance(shape, p):
    """Distance from points p [n, 3] to the primitive (0 inside)."""
    kind, dims = shape
    if kind == "box":
        d = np.maximum(np.abs(p) - np.array(dims) / 2.0, 0.0)
        return np.linalg.norm(d, axis=1)
    r, l = dims
    dr = np.maximum(np.hypot(p[:, 0], p[:, 1]) - r, 0.0)
    dz = np.maximum(np.abs(p[:, 2]) - l / 2.0, 0.0)
    return np.hypot(dr, dz)


def grade(shape, centres, radii, n=20000, seed=0):
    """covered: fraction of the primitive's volume inside the union of the spheres (corners and edges included: the
    sample is volume + the 8 corners / rim points).  protrusion: max over spheres of (distance of the centre from the
    shape + radius) — how far the model reaches outside the shape."""
    centres = np.asarray(centres, dtype=np.float64).reshape(-1, 3)
    radii = np.asarray(radii, dtype=np.float64).reshape(-1)
    rng = np.random.default_rng(seed)
    p = _sample_volume(shape, n, rng)
    kind, dims = shape
    if kind == "box":
        h = np.array(dims) / 2.0
        corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]) * h
    else:
        ang = np.linspace(0, 2 * np.pi, 16, endpoint=False)
        corners = np.concatenate([np.stack([dims[0] * np.cos(ang), dims[0] * np.sin(ang), np.full(16, z)], axis=1)
                                  for z in (-dims[1] / 2, dims[1] / 2)])
    p = np.concatenate([p, corners])
    d = np.linalg.norm(p[:, None, :] - centres[None, :, :], axis=2) - radii[None, :]
    inside = (d <= 1e-12).any(axis=1)
    return {"covered": float(inside.mean()), "protrusion": float((_outside_distance(shape, centres) + radii).max()),
            "n_spheres": int(len(radii))}


def segment_distance(p, a, b):
    """Distance of points p [n, 3] from the segment a-b (the joint offset that bounds a mesh link)."""
    p, a, b = np.asarray(p, float).reshape(-1, 3), np.asarray(a, float), np.asarray(b, float)
    ab = b - a
    t = np.clip(((p - a) @ ab) / max(float(ab @ ab), 1e-30), 0.0, 1.0)
    return np.linalg.norm(p - (a + t[:, None] * ab), axis=1)


def report(link_table, spheres):
    """Per shaped link: its shape, the spheres it owns and — for primitives — the grade of the authored set next to a
    covering fit of the same radius.  Returns a list of dict rows (tools/sphere_model_report.py prints them)."""
    rows = []
    for i, name in enumerate(link_table.names):
        if not link_table.shaped[i]:
            continue
        mine = [(c, r) for n, c, r in spheres if n == name]
        row = {"link": name, "shape": link_table.shape_desc[i], "n_spheres": len(mine),
               "radii": sorted({round(r, 4) for _, r in mine})}
        shape = parse_shape(link_table.shape_desc[i])
        if shape is not None and mine and min(shape[1]) > 0.005:
            c = np.array([c for c, _ in mine])
            r = np.array([r for _, r in mine])
            row["authored"] = grade(shape, c, r)
            fit = fit_primitive(shape, float(r.max()))
            row["fit"] = grade(shape, fit, np.full(len(fit), float(r.max())))
        rows.append(row)
    return rows
