"""Seeded synthetic inputs for the benchmark configurations of BASELINE.json.

Everything here is host-side numpy producing plain arrays; the same arrays feed the CUDA library and
the CPU oracle, so parity tests compare like with like.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

DOF = 8
# URDF:615-759 arm joint limits; base x,y drawn inside the world, theta in [-pi, pi)
ARM_LO = np.array([-1.2, -1.7, -1.8, -2.4, -2.9])
ARM_HI = np.array([1.5, 2.6, 1.8, 2.4, 2.9])
IS_REVOLUTE = np.array([0, 0, 1, 1, 1, 1, 1, 1], dtype=np.int32)   # x, y prismatic (URDF:30-50)
EXTEND_STEP = 0.6        # planner_parameters_robotino.yaml:15
NUM_POINTS = 20          # planner_parameters_robotino.yaml:5  -> 21 states per edge (BS:4473)
NEAR_RADIUS = 1.5        # planner_parameters_robotino.yaml:4
MAX_NEAR = 20            # planner_parameters_robotino.yaml:6


@dataclass
class VoxelWorld:
    bits: np.ndarray      # uint32, bit (k*gy+j)*gx+i
    dims: tuple           # (gx, gy, gz)
    origin: np.ndarray    # float64[3]
    res: float

    def occupancy(self):
        gx, gy, gz = self.dims
        b = np.unpackbits(self.bits.view(np.uint8), bitorder="little")[: gx * gy * gz]
        return b.reshape(gz, gy, gx).astype(bool)


def pack_occupancy(occ_zyx: np.ndarray) -> np.ndarray:
    flat = np.ascontiguousarray(occ_zyx, dtype=np.uint8).reshape(-1)
    pad = (-flat.size) % 32
    if pad:
        flat = np.concatenate([flat, np.zeros(pad, np.uint8)])
    return np.packbits(flat, bitorder="little").view(np.uint32).copy()


def box_world(g: int = 256, res: float = 0.05, n_boxes: int = 64, seed: int = 1234, z_floor: float = 0.0,
              z_low: float = -0.4, dims=None) -> VoxelWorld:
    """W256 / W512 of SURVEY.md §8d: floor slab z < z_floor (the floor the planner inserts,
    squirrel_8dof_planner.cpp:889-902) plus random axis-aligned boxes; grid centred in x,y."""
    gx, gy, gz = dims if dims is not None else (g, g, g)
    origin = np.array([-gx * res / 2.0, -gy * res / 2.0, z_low])
    rng = np.random.default_rng(seed)
    occ = np.zeros((gz, gy, gx), dtype=bool)
    kf = int(round((z_floor - z_low) / res))
    occ[:max(kf, 0)] = True
    half_xy = min(gx, gy) * res / 2.0 - 0.4
    for _ in range(n_boxes):
        c = np.array([rng.uniform(-half_xy, half_xy), rng.uniform(-half_xy, half_xy), rng.uniform(0.0, 2.0)])
        h = rng.uniform(0.1, 0.5, size=3)
        lo = np.floor((c - h - origin) / res).astype(int)
        hi = np.ceil((c + h - origin) / res).astype(int)
        lo = np.maximum(lo, 0)
        hi = np.minimum(hi, [gx, gy, gz])
        if np.all(hi > lo):
            occ[lo[2]:hi[2], lo[1]:hi[1], lo[0]:hi[0]] = True
    return VoxelWorld(pack_occupancy(occ), (gx, gy, gz), origin, res)


def sample_configs(n: int, seed: int, xy_half: float = 5.0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    q = np.empty((n, DOF))
    q[:, 0:2] = rng.uniform(-xy_half, xy_half, size=(n, 2))
    q[:, 2] = rng.uniform(-math.pi, math.pi, size=n)
    q[:, 3:] = rng.uniform(ARM_LO, ARM_HI, size=(n, 5))
    return q


def step_towards(nn: np.ndarray, x_rand: np.ndarray, factor: float = EXTEND_STEP) -> np.ndarray:
    """Vectorised stepTowardsRandSample (BS:5712-5868): same operations in the same order, per row."""
    nn = np.asarray(nn, dtype=np.float64)
    x = np.array(x_rand, dtype=np.float64, copy=True)
    d = x - nn
    sq = d * d
    s_rev = np.zeros(len(nn)); s_pri = np.zeros(len(nn))
    for j in range(DOF):                       # sequential sums, as the reference accumulates them
        if IS_REVOLUTE[j]:
            s_rev = s_rev + sq[:, j]
        else:
            s_pri = s_pri + sq[:, j]
    len_rev, len_pri = np.sqrt(s_rev), np.sqrt(s_pri)
    rev_reached, pri_reached = len_rev < 0.001, len_pri < 0.001
    ext = np.zeros_like(x)
    e_rev = np.zeros(len(nn)); e_pri = np.zeros(len(nn))
    with np.errstate(divide="ignore", invalid="ignore"):
        for j in range(DOF):
            length = len_rev if IS_REVOLUTE[j] else len_pri
            reached = rev_reached if IS_REVOLUTE[j] else pri_reached
            c_step = factor * (d[:, j] / length)
            c_step = np.where(reached, 0.0, c_step)
            ext[:, j] = nn[:, j] + c_step
            if IS_REVOLUTE[j]:
                e_rev = e_rev + c_step * c_step
            else:
                e_pri = e_pri + c_step * c_step
    ext_len_rev = np.where(e_rev == 0.0, 1000.0, np.sqrt(e_rev))
    ext_len_pri = np.where(e_pri == 0.0, 1000.0, np.sqrt(e_pri))
    take_rev, take_pri = ext_len_rev < len_rev, ext_len_pri < len_pri
    for j in range(DOF):
        take = take_rev if IS_REVOLUTE[j] else take_pri
        x[:, j] = np.where(take, ext[:, j], x[:, j])
    return x


def edge_batch(n_edges: int, seed: int = 42, xy_half: float = 5.0):
    """Candidate edges as expandTree makes them (BS:2224-2256): a tree node and a random sample pulled back to one
    extend step (0.6 in the prismatic and in the revolute subspace)."""
    start = sample_configs(n_edges, seed, xy_half)
    rand = sample_configs(n_edges, seed + 1000003, xy_half)
    end = step_towards(start, rand, EXTEND_STEP)
    return start, end


def interpolate_edges(start: np.ndarray, end: np.ndarray, num_points: int = NUM_POINTS) -> np.ndarray:
    """BS:4461-4478: q_k = start + k * ((end - start) / num_points); plain fp64 ops, no fused multiply-add."""
    step = (end - start) / float(num_points)
    k = np.arange(num_points + 1, dtype=np.float64)[None, :, None]
    return (start[:, None, :] + k * step[:, None, :]).reshape(-1, DOF)


def state_batch(n_states: int, seed: int = 42, xy_half: float = 5.0, num_points: int = NUM_POINTS) -> np.ndarray:
    """Config 2: n_states interpolated states = ceil(n/21) edges x 21, truncated."""
    p1 = num_points + 1
    n_edges = (n_states + p1 - 1) // p1
    a, b = edge_batch(n_edges, seed, xy_half)
    return np.ascontiguousarray(interpolate_edges(a, b, num_points)[:n_states])


def random_tree(n_nodes: int, seed: int = 7, cost_seed: int = 9, xy_half: float = 5.0):
    """Config 3: nodes ~ the state distribution, cost ~ U(0, 20)."""
    q = sample_configs(n_nodes, seed, xy_half)
    cost = np.random.default_rng(cost_seed).uniform(0.0, 20.0, size=n_nodes)
    return q, cost
