"""Octomap binary tree (.bt) ingest: the step in front of CollisionChecker::setOcTree (collision_checker.hpp:76-88).

The reference receives an ``octomap::OcTree`` (octomap_msgs::binaryMsgToMap, squirrel_8dof_planner.cpp:872-878 — the
same byte stream as a ``.bt`` file body), inserts a floor patch around the robot (squirrel_8dof_planner.cpp:886-904)
and hands the tree to FCL.  The CUDA path wants a dense bit-packed grid, so this module
  * parses the ``.bt`` stream (format: octomap's OcTree::readBinaryData/readBinaryNode — per inner node two bytes,
    2 bits per child: (bit 2i, bit 2i+1) = (1,0) free leaf, (0,1) occupied leaf, (1,1) inner node, (0,0) unknown;
    depth-first, children of an inner node follow in child order; child i sits at x = i&1, y = (i>>1)&1, z = (i>>2)&1;
    16 levels, key = floor(coord / res) + 32768),
  * expands pruned leaves (a leaf at depth d < 16 covers (2^(16-d))^3 finest voxels),
  * applies the floor patch, and
  * produces a :class:`VoxelWorld` (unknown = free, exactly how FCL treats an OcTree: only occupied leaves collide).
``write_bt`` is the inverse (pruned), used by the tests for round trips.  liboctomap is not needed (and not installed).
"""
from __future__ import annotations

import numpy as np

from .synthetic import VoxelWorld, pack_occupancy

TREE_DEPTH = 16
KEY_ORIGIN = 1 << (TREE_DEPTH - 1)


class BtError(ValueError):
    pass


def parse_bt(data: bytes):
    """-> (res, occupied, free, n_nodes): leaves as arrays [n,4] of (key_x, key_y, key_z, size_in_finest_voxels);
    key = minimum corner of the leaf in finest-voxel key space.  n_nodes counts every node incl. the root."""
    pos = data.find(b"\ndata\n")
    if pos < 0 or not data.startswith(b"# Octomap OcTree binary file"):
        raise BtError("not an octomap .bt stream")
    header = data[:pos].decode("ascii", "replace").split("\n")
    fields = dict(l.split(None, 1) for l in header if l and not l.startswith("#") and " " in l)
    if fields.get("id", "OcTree") != "OcTree":
        raise BtError("unsupported tree id " + fields.get("id", "?"))
    res, size = float(fields["res"]), int(fields["size"])
    body = memoryview(data)[pos + 6:]
    occ, free = [], []
    n_nodes = 1 if size > 0 else 0
    if size > 0:
        # explicit stack instead of recursion: (byte offset is global and advances depth-first)
        off = 0
        stack = [(0, 0, 0, 0)]            # (key_x, key_y, key_z, depth) of inner nodes still to read
        while stack:
            kx, ky, kz, depth = stack.pop()
            if off + 2 > len(body):
                raise BtError("truncated data section")
            bits = body[off] | (body[off + 1] << 8)
            off += 2
            half = 1 << (TREE_DEPTH - depth - 1)
            inner = []
            for i in range(8):
                b0, b1 = (bits >> (2 * i)) & 1, (bits >> (2 * i + 1)) & 1
                if not (b0 or b1):
                    continue
                n_nodes += 1
                cx, cy, cz = kx + (i & 1) * half, ky + ((i >> 1) & 1) * half, kz + ((i >> 2) & 1) * half
                if b0 and b1:
                    if depth + 1 >= TREE_DEPTH:
                        raise BtError("inner node below the finest level")
                    inner.append((cx, cy, cz, depth + 1))
                elif b1:
                    occ.append((cx, cy, cz, half))
                else:
                    free.append((cx, cy, cz, half))
            stack.extend(reversed(inner))   # depth-first, children in ascending order
        if off != len(body):
            raise BtError("%d unread bytes after the tree" % (len(body) - off))
    if size and n_nodes != size:
        raise BtError("header says %d nodes, stream holds %d" % (size, n_nodes))
    as_arr = lambda l: np.array(l, dtype=np.int64).reshape(-1, 4)
    return res, as_arr(occ), as_arr(free), n_nodes


def insert_floor(occ_keys: np.ndarray, res: float, pose_xy, floor_collision_distance: float) -> np.ndarray:
    """squirrel_8dof_planner.cpp:886-904: every finest voxel of the layer just below z = 0 within
    +-int(floorCollisionDistance / res) keys of the robot's (x, y) becomes occupied."""
    kx, ky, kz = (int(np.floor(v / res)) + KEY_ORIGIN for v in (pose_xy[0], pose_xy[1], -0.5 * res))
    n = int(floor_collision_distance / res)
    xs, ys = np.meshgrid(np.arange(kx - n, kx + n + 1), np.arange(ky - n, ky + n + 1), indexing="ij")
    patch = np.stack([xs.ravel(), ys.ravel(), np.full(xs.size, kz), np.ones(xs.size, dtype=np.int64)], axis=1)
    return np.concatenate([occ_keys.reshape(-1, 4), patch]) if len(occ_keys) else patch


def to_voxel_world(res: float, occ_keys: np.ndarray, margin: int = 1, bounds=None) -> VoxelWorld:
    """Dense grid over the bounding box of the occupied leaves (+margin voxels), or over ``bounds`` =
    ((x0,y0,z0),(x1,y1,z1)) in metres.  Pruned leaves are expanded to finest voxels."""
    occ_keys = np.asarray(occ_keys, dtype=np.int64).reshape(-1, 4)
    if bounds is None:
        if len(occ_keys) == 0:
            raise BtError("empty tree: give explicit bounds")
        lo = occ_keys[:, :3].min(0) - margin
        hi = (occ_keys[:, :3] + occ_keys[:, 3:4]).max(0) + margin
    else:
        lo = np.floor(np.asarray(bounds[0]) / res + 1e-9).astype(np.int64) + KEY_ORIGIN   # lattice-aligned bounds stay put
        hi = np.ceil(np.asarray(bounds[1]) / res - 1e-9).astype(np.int64) + KEY_ORIGIN
    dims = (hi - lo).astype(int)
    grid = np.zeros((dims[2], dims[1], dims[0]), dtype=bool)
    for kx, ky, kz, s in occ_keys:
        a = np.maximum(np.array([kx, ky, kz]) - lo, 0)
        b = np.minimum(np.array([kx, ky, kz]) + s - lo, dims)
        if np.all(b > a):
            grid[a[2]:b[2], a[1]:b[1], a[0]:b[0]] = True
    origin = (lo - KEY_ORIGIN).astype(np.float64) * res
    return VoxelWorld(pack_occupancy(grid), (int(dims[0]), int(dims[1]), int(dims[2])), origin, res)


def load_bt(path: str, pose_xy=None, floor_collision_distance: float = 0.0, margin: int = 1, bounds=None) -> VoxelWorld:
    with open(path, "rb") as f:
        res, occ, _free, _n = parse_bt(f.read())
    if pose_xy is not None and floor_collision_distance > 0.0:
        occ = insert_floor(occ, res, pose_xy, floor_collision_distance)
    return to_voxel_world(res, occ, margin, bounds)


def write_bt(world: VoxelWorld) -> bytes:
    """Dense grid -> pruned .bt stream (occupied leaves only; everything else unknown).  The grid origin must lie on
    the voxel lattice (origin / res integral), as every grid made by ``to_voxel_world`` does."""
    res = float(world.res)
    lo = np.rint(np.asarray(world.origin) / res).astype(np.int64) + KEY_ORIGIN
    occ = world.occupancy()                                   # [z, y, x]
    gx, gy, gz = world.dims
    full = np.zeros((1 << TREE_DEPTH,) * 0, dtype=bool)       # placeholder to keep linters quiet
    del full
    n_nodes = [0]
    out = bytearray()

    def region(kx, ky, kz, size):
        """(any, all) occupancy of the cube [k, k+size)^3 in key space."""
        a = np.array([kx, ky, kz]) - lo
        b = a + size
        a_c, b_c = np.maximum(a, 0), np.minimum(b, [gx, gy, gz])
        if np.any(b_c <= a_c):
            return False, False
        sub = occ[a_c[2]:b_c[2], a_c[1]:b_c[1], a_c[0]:b_c[0]]
        inside = bool(np.all(a >= 0) and np.all(b <= [gx, gy, gz]))
        return bool(sub.any()), inside and bool(sub.all())

    def emit(kx, ky, kz, depth):
        half = 1 << (TREE_DEPTH - depth - 1)
        bits, inner = 0, []
        for i in range(8):
            cx, cy, cz = kx + (i & 1) * half, ky + ((i >> 1) & 1) * half, kz + ((i >> 2) & 1) * half
            any_, all_ = region(cx, cy, cz, half)
            if not any_:
                continue
            n_nodes[0] += 1
            if all_ or half == 1:
                bits |= 2 << (2 * i)                          # (bit 2i, bit 2i+1) = (0, 1): occupied leaf
            else:
                bits |= 3 << (2 * i)
                inner.append((cx, cy, cz, depth + 1))
        out.append(bits & 0xFF)
        out.append(bits >> 8)
        for c in inner:
            emit(*c)

    if occ.any():
        n_nodes[0] = 1
        emit(0, 0, 0, 0)
    header = "# Octomap OcTree binary file\n# (written by squirrel_motion_planner_b200.octomap_io)\nid OcTree\nsize %d\nres %.17g\ndata\n" % (
        n_nodes[0], res)
    return header.encode("ascii") + bytes(out)
