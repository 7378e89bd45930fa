"""Robot description for the state-validity hot path (host side, numpy only).

What the reference does at start-up, restated for the new boundary:

* ``CollisionChecker::initializeKDL/createLinksKDL/expandTreeKDL``
  (collision_checker.hpp:176-261) walk the KDL tree depth-first and build a flat
  ``links`` vector: root = identity with z=+0.02 (CC:200-201), the eight named
  planning links keep their KDL joint (CC:216-255), every other link is frozen at
  ``segment.getFrameToTip()`` (CC:257).
* ``initializeFCL`` (CC:263-351) attaches one collision shape per URDF link that
  has one and post-multiplies the collision origin into ``transformToParent``
  (CC:329-336).
* ``findSelfCollisionPairs/parseSRDF`` (CC:353-448) enable every shaped link pair
  that is not listed in the SRDF ``disable_collisions`` table.

This module produces (1) the same flat *link table* from a URDF+SRDF string pair
(`link_table_from_urdf`), which is what the CPU oracle consumes, and (2) the
*compiled* description the CUDA library consumes (`compile_robot`): the frozen
links are folded into rigid groups hanging off the eight planning joints, and
the collision spheres (sphere_model.py) are expressed in their group's frame.

The two consumers use different decompositions on purpose: the oracle walks all
39 links like the reference, the kernels walk 8 chain segments.  Parity between
them is therefore a real check of the folding done here.
"""
from __future__ import annotations

import json
import math
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

DOF = 8
# CC:216-255 — link name -> index into the 8-vector of joint positions
PLANNING_LINKS = {
    "base_x_link": 0, "base_y_link": 1, "base_theta_link": 2, "arm_link1": 3,
    "arm_motor2": 4, "arm_link3": 5, "arm_link4": 6, "arm_link5": 7,
}
ROOT_Z_OFFSET = 0.02      # CC:201
OCTREE_Z_OFFSET = -0.02   # CC:87
JOINT_NONE, JOINT_ROT, JOINT_TRANS = 0, 1, 2

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
DEFAULT_FIXTURE = os.path.join(DATA_DIR, "robotino_plan.json")


# --------------------------------------------------------------------------
# small fp64 frame helpers (set-up time only; the hot path never runs these)
# --------------------------------------------------------------------------
def quat_from_rpy(r, p, y):
    """urdf::Rotation::setFromRPY (urdfdom_headers pose.h) — x,y,z,w."""
    phi, the, psi = r / 2.0, p / 2.0, y / 2.0
    x = math.sin(phi) * math.cos(the) * math.cos(psi) - math.cos(phi) * math.sin(the) * math.sin(psi)
    yy = math.cos(phi) * math.sin(the) * math.cos(psi) + math.sin(phi) * math.cos(the) * math.sin(psi)
    z = math.cos(phi) * math.cos(the) * math.sin(psi) - math.sin(phi) * math.sin(the) * math.cos(psi)
    w = math.cos(phi) * math.cos(the) * math.cos(psi) + math.sin(phi) * math.sin(the) * math.sin(psi)
    n = math.sqrt(x * x + yy * yy + z * z + w * w)
    return x / n, yy / n, z / n, w / n


def rot_from_quat(x, y, z, w):
    """KDL::Rotation::Quaternion (frames.cpp) — row-major 3x3."""
    x2, y2, z2, w2 = x * x, y * y, z * z, w * w
    return np.array([
        [w2 + x2 - y2 - z2, 2 * x * y - 2 * w * z, 2 * x * z + 2 * w * y],
        [2 * x * y + 2 * w * z, w2 - x2 + y2 - z2, 2 * y * z - 2 * w * x],
        [2 * x * z - 2 * w * y, 2 * y * z + 2 * w * x, w2 - x2 - y2 + z2]], dtype=np.float64)


def frame(R=None, p=None):
    T = np.eye(4, dtype=np.float64)
    if R is not None:
        T[:3, :3] = R
    if p is not None:
        T[:3, 3] = p
    return T


def frame_from_xyz_rpy(xyz, rpy):
    return frame(rot_from_quat(*quat_from_rpy(*rpy)), np.asarray(xyz, dtype=np.float64))


def _floats(s, n=3, default=0.0):
    if s is None:
        return [default] * n
    v = [float(t) for t in s.split()]
    assert len(v) == n, s
    return v


# --------------------------------------------------------------------------
# link table  (the oracle's input; mirrors CollisionChecker::links)
# --------------------------------------------------------------------------
@dataclass
class LinkTable:
    names: list
    parent: np.ndarray            # int32 [L]; -1 for the root
    joint_type: np.ndarray        # int32 [L]; JOINT_* for planning links, NONE otherwise
    qidx: np.ndarray              # int32 [L]; index into q[8] or -1
    joint_origin: np.ndarray      # f64 [L,3]   KDL Joint origin (planning links)
    joint_axis: np.ndarray        # f64 [L,3]   KDL Joint axis, normalised
    to_parent: np.ndarray         # f64 [L,12]  rows of R then p — transformToParent (frozen links)
    frame_to_tip: np.ndarray      # f64 [L,12]  KDL segment.getFrameToTip() = URDF parent->joint origin frame F
    shaped: np.ndarray            # int32 [L]   1 if the link owns a collision shape (CC:324-343)
    shape_desc: list              # human readable shape per link ('' if none)
    disabled_pairs: list          # [(name, name)] from the SRDF (CC:395-419)
    octomap_disabled: list        # links removed from the map check by SRDF "octomap" lines (CC:408-416)
    joint_limits: np.ndarray      # f64 [8,2] lower/upper (NaN for continuous)
    extras: dict = field(default_factory=dict)

    @property
    def n(self):
        return len(self.names)

    def index(self, name):
        return self.names.index(name)

    def self_pair_matrix(self):
        """collisionMatrix of CC:373-385: True = pair is checked (both shaped, not disabled)."""
        L = self.n
        m = np.ones((L, L), dtype=bool)
        for a, b in self.disabled_pairs:
            if a in self.names and b in self.names:
                i, j = self.index(a), self.index(b)
                m[i, j] = m[j, i] = False
        sh = self.shaped.astype(bool)
        m &= sh[:, None] & sh[None, :]
        np.fill_diagonal(m, False)
        return m

    def to_json(self):
        return {
            "names": self.names, "parent": self.parent.tolist(), "joint_type": self.joint_type.tolist(),
            "qidx": self.qidx.tolist(), "joint_origin": self.joint_origin.tolist(),
            "joint_axis": self.joint_axis.tolist(), "to_parent": self.to_parent.tolist(),
            "frame_to_tip": self.frame_to_tip.tolist(), "shaped": self.shaped.tolist(), "shape_desc": self.shape_desc,
            "disabled_pairs": [list(p) for p in self.disabled_pairs],
            "octomap_disabled": self.octomap_disabled,
            "joint_limits": [[None if math.isnan(v) else v for v in row] for row in self.joint_limits.tolist()],
            "extras": self.extras,
        }

    @staticmethod
    def from_json(d):
        lim = np.array([[math.nan if v is None else v for v in row] for row in d["joint_limits"]], dtype=np.float64)
        return LinkTable(
            names=list(d["names"]), parent=np.array(d["parent"], dtype=np.int32),
            joint_type=np.array(d["joint_type"], dtype=np.int32), qidx=np.array(d["qidx"], dtype=np.int32),
            joint_origin=np.array(d["joint_origin"], dtype=np.float64),
            joint_axis=np.array(d["joint_axis"], dtype=np.float64),
            to_parent=np.array(d["to_parent"], dtype=np.float64),
            frame_to_tip=np.array(d["frame_to_tip"], dtype=np.float64), shaped=np.array(d["shaped"], dtype=np.int32),
            shape_desc=list(d["shape_desc"]), disabled_pairs=[tuple(p) for p in d["disabled_pairs"]],
            octomap_disabled=list(d.get("octomap_disabled", [])), joint_limits=lim, extras=d.get("extras", {}))


def parse_srdf_disabled(srdf_xml: str):
    """CC:395-448 restated: line-based scan for ``<disable_collisions link1= link2=``."""
    pairs, octo = [], []
    for line in srdf_xml.splitlines():
        if "<disable_collisions" not in line:
            continue
        def grab(tag):
            i = line.find(tag) + len(tag) + 1
            j = line.find('"', i)
            return line[i:j]
        a, b = grab("link1="), grab("link2=")
        if a != "octomap" and b != "octomap":
            pairs.append((a, b))
        elif not (a == "octomap" and b == "octomap"):
            octo.append(b if a == "octomap" else a)
    return pairs, octo


def link_table_from_urdf(urdf_xml: str, srdf_xml: str) -> LinkTable:
    """Flat DFS link table, as CollisionChecker builds it (CC:176-261, 263-351).

    Children are visited in joint-name order: urdfdom fills ``child_links`` while
    iterating its name-sorted joint map, kdl_parser adds segments in that order,
    and CC:211 walks ``segment.children`` in insertion order.  The order only
    affects early exit, never a verdict.
    """
    root = ET.fromstring(urdf_xml)
    links = {l.get("name"): l for l in root.findall("link")}
    joints = {}
    children_of = {}
    child_names = set()
    for j in root.findall("joint"):
        name = j.get("name")
        parent, child = j.find("parent").get("link"), j.find("child").get("link")
        o = j.find("origin")
        xyz = _floats(o.get("xyz") if o is not None else None)
        rpy = _floats(o.get("rpy") if o is not None else None)
        a = j.find("axis")
        axis = _floats(a.get("xyz") if a is not None else None) if a is not None else [1.0, 0.0, 0.0]
        lim = j.find("limit")
        joints[child] = dict(name=name, type=j.get("type"), parent=parent, xyz=xyz, rpy=rpy, axis=axis,
                             lower=float(lim.get("lower")) if lim is not None and lim.get("lower") else math.nan,
                             upper=float(lim.get("upper")) if lim is not None and lim.get("upper") else math.nan)
        children_of.setdefault(parent, []).append((name, child))
        child_names.add(child)
    roots = [n for n in links if n not in child_names]
    assert len(roots) == 1, roots

    names, parent, jtype, qidx, jorig, jaxis, to_parent, shaped, sdesc = [], [], [], [], [], [], [], [], []
    ftip = []
    limits = np.full((DOF, 2), math.nan)

    def collision_of(name):
        c = links[name].find("collision")
        if c is None or c.find("geometry") is None:
            return None
        g = c.find("geometry")[0]
        ok, desc = False, ""
        if g.tag == "box":
            d = _floats(g.get("size"))
            ok, desc = all(v > 0 for v in d), "box %g %g %g" % tuple(d)
        elif g.tag == "cylinder":
            r, l = float(g.get("radius")), float(g.get("length"))
            ok, desc = r > 0 and l > 0, "cylinder r=%g l=%g" % (r, l)
        elif g.tag == "sphere":
            r = float(g.get("radius"))
            ok, desc = r > 0, "sphere r=%g" % r
        elif g.tag == "mesh":
            ok, desc = True, "mesh " + g.get("filename").rsplit("/", 1)[-1]
        if not ok:
            return None
        o = c.find("origin")
        xyz = _floats(o.get("xyz") if o is not None else None)
        rpy = _floats(o.get("rpy") if o is not None else None)
        return frame_from_xyz_rpy(xyz, rpy), desc

    def add(name, par):
        idx = len(names)
        names.append(name)
        parent.append(par)
        T = np.eye(4)
        F = np.eye(4)
        jt, qi, jo, ja = JOINT_NONE, -1, np.zeros(3), np.zeros(3)
        if par >= 0:
            j = joints[name]
            F = frame_from_xyz_rpy(j["xyz"], j["rpy"])
            if name in PLANNING_LINKS:
                qi = PLANNING_LINKS[name]
                # kdl_parser toKdl(joint): Joint(name, F.p, F.M*axis, RotAxis|TransAxis); KDL normalises the axis
                ax = F[:3, :3] @ np.asarray(j["axis"], dtype=np.float64)
                ja = ax / np.linalg.norm(ax)
                jo = F[:3, 3].copy()
                if j["type"] == "prismatic":
                    jt = JOINT_TRANS
                elif j["type"] in ("revolute", "continuous"):
                    jt = JOINT_ROT
                else:
                    raise ValueError("planning link %s hangs on a %s joint" % (name, j["type"]))
                limits[qi] = (j["lower"], j["upper"])
            else:
                T = F            # getFrameToTip(): the joint frozen at 0 (CC:257)
        col = collision_of(name)
        if col is not None:
            T = T @ col[0]       # CC:336 (only read back for frozen links, CC:527-530)
        jtype.append(jt); qidx.append(qi); jorig.append(jo); jaxis.append(ja)
        to_parent.append(np.concatenate([T[:3, :3].reshape(9), T[:3, 3]]))
        ftip.append(np.concatenate([F[:3, :3].reshape(9), F[:3, 3]]))
        shaped.append(1 if col is not None else 0)
        sdesc.append(col[1] if col is not None else "")
        for _, child in sorted(children_of.get(name, [])):
            add(child, idx)

    add(roots[0], -1)
    pairs, octo = parse_srdf_disabled(srdf_xml)
    return LinkTable(names, np.array(parent, np.int32), np.array(jtype, np.int32), np.array(qidx, np.int32),
                     np.array(jorig), np.array(jaxis), np.array(to_parent), np.array(ftip), np.array(shaped, np.int32), sdesc,
                     pairs, octo, limits)


def rigid_group_of_links(lt: LinkTable) -> np.ndarray:
    """Index of the planning link each link is rigid to (-1: the root) — frozen joints do not move (CC:257)."""
    grp = np.full(lt.n, -1, np.int32)
    for i in range(lt.n):
        p = int(lt.parent[i])
        if p >= 0:
            grp[i] = i if lt.joint_type[i] != JOINT_NONE else grp[p]
    return grp


def disable_rigid_pairs(lt: LinkTable):
    """Deliberate deviation from CC:353-393, applied when the fixture is generated: shaped link pairs that are rigid to
    each other (same planning link) are taken out of the self-collision matrix.  Their relative pose never changes, so
    they cannot distinguish one configuration from another — but the reference's SRDF predates the box that replaced
    the front-shell mesh (URDF:324-328) and leaves pairs enabled whose URDF shapes overlap permanently (the hokuyo
    cylinder lies inside that box): with them every configuration would be "in self collision".  Returns the pairs
    that were enabled before (recorded in the fixture's extras)."""
    grp = rigid_group_of_links(lt)
    m = lt.self_pair_matrix()
    added = []
    for a in range(lt.n):
        for b in range(a + 1, lt.n):
            if m[a, b] and grp[a] == grp[b]:
                added.append((lt.names[a], lt.names[b]))
    lt.disabled_pairs = list(lt.disabled_pairs) + added
    lt.extras = dict(lt.extras, rigid_pairs_disabled=[list(p) for p in added])
    return added


def load_fixture(path: str = DEFAULT_FIXTURE):
    """Link table + authored sphere model + reference fixtures, as committed under data/."""
    with open(path) as f:
        d = json.load(f)
    return LinkTable.from_json(d["link_table"]), d


# --------------------------------------------------------------------------
# compiled description  (the CUDA library's input: sqmp_robot_desc)
# --------------------------------------------------------------------------
@dataclass
class CompiledRobot:
    """Flat arrays matching ``sqmp_robot_desc`` in include/sqmp_b200.h."""
    n_seg: int
    seg_type: np.ndarray     # i32 [n_seg]   JOINT_ROT / JOINT_TRANS
    seg_qidx: np.ndarray     # i32 [n_seg]
    seg_pre: np.ndarray      # f64 [n_seg,12] fixed frame applied before the joint motion (R rows, p)
    seg_axis: np.ndarray     # f64 [n_seg,3]
    root_offset: np.ndarray  # f64 [3]  CollisionChecker's root translation (CC:201); not part of compute_FK
    ee_tip: np.ndarray       # f64 [12] fixed frame from the last segment to the FK tip link
    n_spheres: int
    sph_seg: np.ndarray      # i32 [n]  segment whose frame carries the sphere (-1 = root)
    sph_link: np.ndarray     # i32 [n]  link-table index (for the per-request link disable mask)
    sph_center: np.ndarray   # f64 [n,3] in the segment frame
    sph_radius: np.ndarray   # f64 [n]
    link_pair_mask: np.ndarray  # u64 [64]  bit j of row i: links i and j are checked against each other (CC:353-393)
    self_const_collision: int  # 1 if an enabled pair inside one rigid group always collides
    link_names: list
    ee_link: str


def _f12(T):
    return np.concatenate([T[:3, :3].reshape(9), T[:3, 3]])


def compile_robot(lt: LinkTable, spheres, ee_link: str = "hand_wrist_link") -> CompiledRobot:
    """Fold the 39-link table into a chain of planning segments + rigid sphere groups.

    ``spheres``: list of (link_name, (x,y,z), r) in the link's collision frame
    (= ``Link::transform`` of the reference, i.e. after the collision origin).
    World frame of segment k:  T_k = T_{k-1} * seg_pre[k] * Motion(seg_axis[k], q[seg_qidx[k]]),
    T_{-1} = Frame(root_offset) for the collision path and identity for compute_FK (KM:278).
    """
    L = lt.n
    seg_of_link = np.full(L, -2, np.int32)       # segment whose frame the link is rigid to
    rel = [None] * L                             # collision frame of the link in that segment's frame
    rel_fk = [None] * L                          # same, from KDL frame_to_tip only (no collision origins)
    seg_type, seg_qidx, seg_pre, seg_axis = [], [], [], []
    for i in range(L):
        p = int(lt.parent[i])
        if p < 0:
            seg_of_link[i] = -1
            rel[i] = np.eye(4); rel_fk[i] = np.eye(4)
            continue
        if lt.joint_type[i] != JOINT_NONE:
            k = len(seg_type)
            if seg_of_link[p] != k - 1:
                raise ValueError("planning joints must form one serial chain (link %s)" % lt.names[i])
            ft = lt.frame_to_tip[i]
            if not np.allclose(ft[:9].reshape(3, 3), np.eye(3), atol=0.0):
                # CC:528 uses joint->pose(q) (no f_tip); KDL's chain solver uses segment.pose(q).  They agree
                # only for a joint origin without rotation — true for all eight planning joints of this robot.
                raise ValueError("planning joint with rotated origin is outside the reference's own assumption")
            if not np.array_equal(rel[p], rel_fk[p]):
                raise ValueError("collision origin of %s would leak into the kinematic chain" % lt.names[p])
            seg_pre.append(rel[p] @ frame(p=lt.joint_origin[i]))
            seg_type.append(int(lt.joint_type[i])); seg_qidx.append(int(lt.qidx[i]))
            seg_axis.append(lt.joint_axis[i].copy())
            seg_of_link[i] = k
            rel[i] = np.eye(4); rel_fk[i] = np.eye(4)
        else:
            tp, ft = lt.to_parent[i], lt.frame_to_tip[i]
            seg_of_link[i] = seg_of_link[p]
            rel[i] = rel[p] @ frame(tp[:9].reshape(3, 3), tp[9:])
            rel_fk[i] = rel_fk[p] @ frame(ft[:9].reshape(3, 3), ft[9:])
    n_seg = len(seg_type)
    ee = lt.index(ee_link)
    if seg_of_link[ee] != n_seg - 1:
        raise ValueError("FK tip link must be rigid to the last planning segment")

    sph_seg, sph_link, sph_c, sph_r = [], [], [], []
    for name, c, r in spheres:
        i = lt.index(name)
        if not lt.shaped[i]:
            raise ValueError("sphere on link without a collision shape: " + name)
        cc = rel[i] @ np.array([c[0], c[1], c[2], 1.0])
        sph_seg.append(int(seg_of_link[i])); sph_link.append(i); sph_c.append(cc[:3]); sph_r.append(float(r))
    order = np.lexsort((np.arange(len(sph_seg)), np.array(sph_seg)))   # group by segment, stable
    sph_seg = np.array(sph_seg, np.int32)[order]; sph_link = np.array(sph_link, np.int32)[order]
    sph_c = np.array(sph_c, np.float64)[order]; sph_r = np.array(sph_r, np.float64)[order]

    # selfCollisionPairs at link level (CC:353-393); the library collides the spheres of an enabled link pair when the
    # two links hang on different segments.  A pair inside one rigid group never changes: if it overlaps, it always does.
    m = lt.self_pair_matrix()
    rows = np.zeros(64, dtype=np.uint64)
    for a in range(L):
        for b in range(L):
            if m[a, b]:
                rows[a] |= np.uint64(1) << np.uint64(b)
    const_hit = 0
    by_link = {}
    for i in range(len(sph_r)):
        by_link.setdefault(int(sph_link[i]), []).append(i)
    for a in by_link:
        for b in by_link:
            if a < b and m[a, b] and seg_of_link[a] == seg_of_link[b]:
                ia, ib = np.array(by_link[a]), np.array(by_link[b])
                d = np.linalg.norm(sph_c[ia][:, None, :] - sph_c[ib][None, :, :], axis=2) - (sph_r[ia][:, None] + sph_r[ib][None, :])
                if d.min() <= 0.0:
                    const_hit = 1
    return CompiledRobot(
        n_seg, np.array(seg_type, np.int32), np.array(seg_qidx, np.int32),
        np.array([_f12(T) for T in seg_pre]), np.array(seg_axis),
        np.array([0.0, 0.0, ROOT_Z_OFFSET]), _f12(rel_fk[ee]),
        len(sph_r), sph_seg, sph_link, sph_c, sph_r, rows, const_hit, list(lt.names), ee_link)


def link_mask(lt: LinkTable, disabled_names) -> int:
    """Bit i set = link i is *excluded* from the map check (setDisabledLinkMapCollisions, CC:90-102)."""
    mask = 0
    for n in list(disabled_names) + list(lt.octomap_disabled):
        if n in lt.names:
            mask |= 1 << lt.index(n)
    return mask
