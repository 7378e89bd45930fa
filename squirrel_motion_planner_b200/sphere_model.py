"""Collision-sphere model of the SQUIRREL Robotino (built here; the reference has none).

The reference collides FCL shapes (CollisionChecker::initializeFCL, collision_checker.hpp:263-351): six URDF primitives
and 21 STL meshes that live in the external ``robotino_description`` package and are not in the repository
(URDF:110-115, 310-315, ...).  The north star replaces them with spheres.  The model is CONSERVATIVE:

  * every one of the six primitives (front shell box URDF:323-330, neck box 355-360, neck-pan cylinder 385-390, head
    cylinder 430-435, hokuyo cylinder 546-551, kinect box 1369-1374) is covered completely by its spheres, which reach
    at most ``P_PRIMITIVE`` outside it (``P_BOTTOM`` below: the floor slab the planner inserts, squirrel_8dof_planner.cpp:
    889-902, lies 0.04 m under the robot's lowest shape and must stay clear of it);
  * every mesh link gets a declared stand-in primitive (``MESH_STANDINS``: capsules along the joint offsets that bound the
    link, URDF:587-759, 843-846; a cylinder for the drum of the mobile base) that is covered the same way with the
    inflation stated next to it.

``build_model`` runs the covering fit of ``sphere_fit.py``; its output is frozen in ``data/robotino_spheres.json`` (the
fit needs scipy and a few seconds; the file is what the library, the oracle and the tests load) and
``tests/test_host_logic.py`` re-grades the frozen model: coverage 1.0 of every link shape, protrusion <= one 0.05 m voxel.
Each entry is ``(link, centre xyz in the link's collision frame, radius)``; every shaped link of the reference owns at
least one sphere so that the per-link disable mask (CC:90-102) and the SRDF pair table (CC:353-393) keep their meaning.
"""
from __future__ import annotations

import json
import os

P_PRIMITIVE = 0.035   # how far the spheres of a URDF primitive may reach outside it (m); one voxel is 0.05
P_BOTTOM = 0.030      # ... below its -z face (floor clearance of the lowest shape is 0.04)
P_ARM = 0.010         # arm / body stand-in capsules
P_HAND = 0.005        # wrist, palm and finger stand-ins

FROZEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "robotino_spheres.json")


def _cap(a, b, R):
    return ("capsule", (tuple(a), tuple(b), R), (0.0, 0.0, 0.0))


def _ball(c, R):
    return ("capsule", (tuple(c), tuple(c), R), (0.0, 0.0, 0.0))


# Stand-ins for the 21 mesh links, in each link's collision frame (= Link::transform of the reference).  (shape, inflation)
MESH_STANDINS = {
    # mobile base: drum of the Robotino platform (diameter 0.45 m), 2 cm to 15 cm above its frame — the published folded
    # arm pose (parameters.yaml:37) rests the wrist at that height; the link frame sits 1 cm in front of the platform
    # axis (URDF base_body_joint)
    "base_body_link": (("cylinder", (0.225, 0.13), (-0.01, 0.0, 0.085)), P_PRIMITIVE),
    # shell: rear column that carries neck and head
    "shell_base_link": (_cap((-0.21, 0.0, 0.30), (-0.21, 0.0, 0.60), 0.09), P_ARM),
    "door_link": (_cap((-0.05, -0.21, 0.35), (-0.05, -0.21, 0.50), 0.04), P_ARM),
    # arm: one capsule per link along the offset to the next joint (URDF:587-759)
    "arm_base_link": (_cap((0.0, 0.0, -0.05), (0.0, 0.0, -0.01), 0.05), P_ARM),
    "arm_link1": (_cap((0.0, 0.0, 0.04), (0.0, 0.0, 0.23), 0.055), P_ARM),          # joint2 at (0, 0, 0.25)
    "arm_link2": (("union", (_cap((0.0, -0.05, 0.0), (0.0, 0.0, 0.04), 0.045),      # motor2 sits at (0, -0.098, -0.027)
                             _cap((0.0, 0.0, 0.04), (0.0, 0.0, 0.125), 0.045)), (0.0, 0.0, 0.0)), P_ARM),   # joint3 at (0, 0, 0.158)
    "arm_link3": (_cap((0.0, -0.015, 0.02), (0.0, -0.05, 0.06), 0.045), P_ARM),     # joint4 at (0, -0.057, 0.065)
    "arm_link4": (_cap((0.0, -0.01, 0.015), (0.0, -0.02, 0.025), 0.045), P_ARM),    # joint5 at (0, -0.027, 0.035)
    # wrist + hand (rigid to arm_link5; wrist y axis = arm_link5 z axis; the palm lies 0.15 m along it, URDF:843-846)
    "hand_wrist_link": (_cap((0.0, 0.07, 0.0), (-0.01, 0.13, 0.0), 0.04), P_HAND),
    "hand_base_link": (_cap((0.0, 0.02, 0.0), (0.0, 0.05, 0.0), 0.025), P_HAND),
    "hand_middle_finger_lower_link": (_cap((0.01, 0.0, 0.0), (0.04, 0.0, 0.0), 0.01), P_HAND),
    "hand_middle_finger_upper_link": (_cap((0.005, 0.0, 0.0), (0.025, 0.0, 0.0), 0.01), P_HAND),
    "hand_left_crank": (_ball((0.0, 0.0, 0.0), 0.008), P_HAND),
    "hand_left_coupler": (_ball((0.008, 0.01, 0.03), 0.01), P_HAND),
    "hand_left_finger_lower_link": (_cap((0.01, 0.0, 0.0), (0.04, 0.0, 0.0), 0.01), P_HAND),
    "hand_left_finger_upper_link": (_cap((0.01, 0.0, 0.0), (0.04, 0.0, 0.0), 0.01), P_HAND),
    "hand_right_crank": (_ball((0.0, 0.0, 0.0), 0.008), P_HAND),
    "hand_right_coupler": (_ball((-0.008, 0.01, 0.03), 0.01), P_HAND),
    "hand_right_finger_lower_link": (_cap((0.01, 0.0, 0.0), (0.04, 0.0, 0.0), 0.01), P_HAND),
    "hand_right_finger_upper_link": (_cap((0.01, 0.0, 0.0), (0.04, 0.0, 0.0), 0.01), P_HAND),
    "hand_cableCanal_link": (_ball((0.0, 0.0, 0.0), 0.025), P_HAND),
}


def link_shapes(link_table):
    """{link: (shape, source)} for every shaped link: the URDF primitive where the reference has one, else the
    declared stand-in."""
    from . import sphere_fit as sf

    out = {}
    for i, name in enumerate(link_table.names):
        if not link_table.shaped[i]:
            continue
        shape = sf.parse_shape(link_table.shape_desc[i])
        if shape is not None:
            out[name] = (shape, "URDF " + link_table.shape_desc[i])
        elif name in MESH_STANDINS:
            out[name] = (MESH_STANDINS[name][0], "stand-in for " + link_table.shape_desc[i])
        else:
            raise KeyError("no stand-in declared for mesh link " + name)
    return out


def inflation_of(link_table, name):
    return MESH_STANDINS[name][1] if name in MESH_STANDINS else P_PRIMITIVE


def build_model(link_table, progress=None):
    """Run the covering fit for every shaped link -> [(link, (x, y, z), r)] (deterministic)."""
    from . import sphere_fit as sf

    spheres = []
    for name, (shape, source) in link_shapes(link_table).items():
        p = inflation_of(link_table, name)
        c, r = sf.cover_shape(shape, p, p_bottom=min(p, P_BOTTOM))
        if progress:
            progress("%-32s %-44s %4d spheres" % (name, source[:44], len(r)))
        spheres += [(name, (round(float(v[0]), 6), round(float(v[1]), 6), round(float(v[2]), 6)), round(float(rr), 6))
                    for v, rr in zip(c, r)]
    return spheres


def spheres_for(robot: str = "robotino"):
    """The frozen model (tools/gen_robot_fixture.py --refit rebuilds it)."""
    if robot != "robotino":
        raise KeyError(robot)
    with open(FROZEN) as f:
        d = json.load(f)
    return [(n, tuple(float(v) for v in c), float(r)) for n, c, r in d["spheres"]]
