"""Octomap .bt ingest: parser / writer round trips, the reference's shipped maps, the floor patch."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from squirrel_motion_planner_b200 import octomap_io as oio, synthetic as syn

ROOMS = "/root/reference/squirrel_8dof_planner/config/"


def test_round_trip_synthetic_world():
    w = syn.box_world(g=64, n_boxes=10, seed=3, dims=(64, 48, 32))     # origin on the voxel lattice
    data = oio.write_bt(w)
    res, occ, free, n = oio.parse_bt(data)
    assert res == w.res and len(free) == 0 and n == int(dict(l.split(None, 1) for l in data[:200].decode("ascii", "replace").split("\n")[2:4])["size"])
    assert (occ[:, 3] ** 3).sum() == w.occupancy().sum()                # pruned leaves cover exactly the occupied voxels
    assert (occ[:, 3] > 1).any()                                        # and pruning did happen (floor slab)
    back = oio.to_voxel_world(res, occ, margin=0, bounds=(w.origin, w.origin + np.array(w.dims) * w.res))
    assert back.dims == w.dims and np.array_equal(back.bits, w.bits) and np.allclose(back.origin, w.origin)


def test_malformed_streams_are_rejected():
    w = syn.box_world(g=32, n_boxes=3, seed=4)
    data = oio.write_bt(w)
    for bad in (b"not a tree", data[:-2], data + b"\0\0", data.replace(b"id OcTree", b"id ColorOcTree")):
        with pytest.raises(oio.BtError):
            oio.parse_bt(bad)


def test_reference_rooms_golden():
    """tests/golden/rooms_bt.npz = room3/4/5.bt of the reference parsed in the build container.  Counts match what
    the files' own headers declare (node count) and SURVEY.md §8c; writing the dense grid back reproduces the tree."""
    g = np.load(os.path.join(GOLDEN, "rooms_bt.npz"))
    want = {"room3": (2687, 1705, 0, 5548), "room4": (1141, 580, 0, 4192), "room5": (1351, 740, 0, 4408)}
    for name, counts in want.items():
        assert tuple(g[name + "_counts"]) == counts
        w = syn.VoxelWorld(g[name + "_bits"], tuple(int(v) for v in g[name + "_dims"]), g[name + "_origin"], float(g[name + "_res"]))
        assert w.occupancy().sum() == counts[3]
        res, occ, free, n = oio.parse_bt(oio.write_bt(w))
        assert (n, len(occ), len(free), int((occ[:, 3] ** 3).sum())) == counts
        if os.path.isdir(ROOMS):                                        # build container: the real file, byte stream in
            res2, occ2, free2, n2 = oio.parse_bt(open(ROOMS + name + ".bt", "rb").read())
            assert (n2, len(occ2), len(free2)) == counts[:3]
            assert np.array_equal(oio.to_voxel_world(res2, occ2).bits, w.bits)


def test_floor_patch_matches_the_planner_node():
    """squirrel_8dof_planner.cpp:886-904: keys within +-int(d/res) of the robot, in the layer just below z = 0."""
    res = 0.05
    occ = oio.insert_floor(np.zeros((0, 4), dtype=np.int64), res, (1.23, -0.4), 2.0)
    n = int(2.0 / res)
    assert len(occ) == (2 * n + 1) ** 2 and set(occ[:, 2]) == {oio.KEY_ORIGIN - 1} and set(occ[:, 3]) == {1}
    w = oio.to_voxel_world(res, occ, margin=0)
    assert w.dims == (2 * n + 1, 2 * n + 1, 1) and abs(w.origin[2] + res) < 1e-12
    cx = (int(np.floor(1.23 / res)) - n) * res
    assert abs(w.origin[0] - cx) < 1e-12


def test_cpp_reader_agrees_with_python(tmp_path):
    """sqmp_shim::read_bt (the C++ side of setOcTree for .bt / binary octomap messages) == octomap_io, bit for bit."""
    import subprocess
    from conftest import ROOT
    shim = os.path.join(ROOT, "squirrel_motion_planner_b200", "shim")
    subprocess.run(["make", "-s", "-C", shim], check=True)
    g = np.load(os.path.join(GOLDEN, "rooms_bt.npz"))
    for name, floor in (("room3", None), ("room5", (0.7, -1.3, 1.5))):
        w = syn.VoxelWorld(g[name + "_bits"], tuple(int(v) for v in g[name + "_dims"]), g[name + "_origin"], float(g[name + "_res"]))
        bt = tmp_path / (name + ".bt")
        bt.write_bytes(oio.write_bt(w))
        args = [os.path.join(shim, "test_shim"), "--bt", str(bt), str(tmp_path / name)]
        if floor:
            args += ["%r" % v for v in floor]
        subprocess.run(args, check=True)
        want = oio.load_bt(str(bt), floor[:2] if floor else None, floor[2] if floor else 0.0)
        meta = open(str(tmp_path / name) + ".meta.txt").read().split()
        assert tuple(int(v) for v in meta[:3]) == want.dims
        assert np.allclose([float(v) for v in meta[3:6]], want.origin, atol=1e-12) and float(meta[6]) == want.res
        assert np.array_equal(np.fromfile(str(tmp_path / name) + ".bits.u32", dtype=np.uint32), want.bits)
