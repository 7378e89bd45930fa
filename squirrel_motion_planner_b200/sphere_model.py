"""Collision-sphere model of the SQUIRREL Robotino (authored here; the reference has none).

The reference collides FCL shapes — 21 of the 27 are STL meshes that live in the
external ``robotino_description`` package and are not in the repository
(URDF:110-115, 310-315, ...).  The north star replaces them with
spheres.  Each entry is ``(link, centre xyz in the link's collision frame, radius)``;
every one of the 27 shaped links of the reference (CC:263-351) owns at least one
sphere so that the per-link disable mask (CC:90-102) and the SRDF pair table
(CC:353-393) keep their meaning.

Sources for the dimensions: the six URDF primitives (front shell box URDF:323-330,
neck box 355-360, neck-pan cylinder 385-390, head cylinder 430-435, hokuyo
cylinder 546-551, kinect box 1369-1374) and, for mesh links, the joint offsets
that bound each link (URDF:587-759, 843-846 and the hand joints).  Radii come from
a small set so the voxel kernels see few distinct scan extents.  The lowest
sphere keeps >= one 0.05 m voxel of slack above the floor slab the planner
inserts (squirrel_8dof_planner.cpp:889-902), like the real body does.
"""

# radius classes (metres)
R_FINGER, R_WRIST, R_ARM, R_ARM_BIG, R_BODY, R_SHELL, R_HEAD = 0.015, 0.045, 0.055, 0.065, 0.06, 0.10, 0.26


def _ring(link, n, radius, z, r, phase=0.0):
    import math
    return [(link, (radius * math.cos(phase + 2 * math.pi * k / n), radius * math.sin(phase + 2 * math.pi * k / n), z), r)
            for k in range(n)]


ROBOTINO_SPHERES = (
    # ---- mobile base (rigid to base_theta_link) --------------------------------
    # RobotinoBody mesh: low drum of radius ~0.22 m under the arm socket (z = 0.25): ring + hub
    _ring("base_body_link", 8, 0.17, 0.08, R_BODY) + _ring("base_body_link", 4, 0.07, 0.08, R_BODY, 0.785398)
    # base_shell_collision mesh: rear column that carries neck and head
    + [("shell_base_link", (-0.21, 0.0, 0.30), R_SHELL),
       ("shell_base_link", (-0.21, 0.0, 0.45), R_SHELL),
       ("shell_base_link", (-0.21, 0.0, 0.60), R_SHELL)]
    # front shell box 0.55 x 0.55 x 0.17 (frame already at the box centre): the four corners the drum misses
    + [("shell_base_link_front", (0.21, 0.21, -0.02), R_BODY),
       ("shell_base_link_front", (0.21, -0.21, -0.02), R_BODY),
       ("shell_base_link_front", (-0.21, 0.21, -0.02), R_BODY),
       ("shell_base_link_front", (-0.21, -0.21, -0.02), R_BODY)]
    # 1 mm dummy box
    + [("base_neck_link", (0.0, 0.0, 0.0), 0.001)]
    # neck-pan cylinder r=0.2 l=0.08: flat ring + hub
    + _ring("neck_pan_link", 6, 0.14, 0.0, 0.06) + [("neck_pan_link", (0.0, 0.0, 0.0), 0.06)]
    # head cylinder r=0.23 l=0.25 (frame at the cylinder centre)
    + [("head_link", (0.0, 0.0, 0.0), R_HEAD)]
    # door mesh on the side of the shell
    + [("door_link", (-0.05, -0.21, 0.35), 0.05), ("door_link", (-0.05, -0.21, 0.50), 0.05)]
    # hokuyo cylinder r=0.03 l=0.07
    + [("hokuyo_link", (0.0, 0.0, 0.0), R_WRIST)]
    # arm socket (link0 mesh)
    + [("arm_base_link", (0.0, 0.0, -0.03), 0.06)]
    # kinect box 0.035 x 0.18 x 0.02
    + [("kinect_link", (0.0, 0.05, 0.0), 0.04), ("kinect_link", (0.0, -0.05, 0.0), 0.04)]
    # ---- arm ------------------------------------------------------------------
    # link1: joint1 -> joint2 at (0, 0, 0.25)
    + [("arm_link1", (0.0, 0.0, 0.05), R_ARM_BIG), ("arm_link1", (0.0, 0.0, 0.14), R_ARM_BIG),
       ("arm_link1", (0.0, 0.0, 0.23), R_ARM_BIG)]
    # link2 (offset 0.098 in y from motor2): -> joint3 at (0, 0, 0.158)
    + [("arm_link2", (0.0, -0.045, 0.0), R_ARM), ("arm_link2", (0.0, 0.0, 0.045), R_ARM),
       ("arm_link2", (0.0, 0.0, 0.125), R_ARM)]
    # link3: -> joint4 at (0, -0.057, 0.065)
    + [("arm_link3", (0.0, -0.015, 0.02), R_ARM), ("arm_link3", (0.0, -0.050, 0.06), R_ARM)]
    # link4: -> joint5 at (0, -0.027, 0.035)
    + [("arm_link4", (0.0, -0.015, 0.02), R_ARM)]
    # ---- wrist + hand (rigid to arm_link5; wrist y axis = arm_link5 z axis) ----
    + [("hand_wrist_link", (0.0, 0.075, 0.0), R_WRIST), ("hand_wrist_link", (-0.01, 0.125, 0.0), R_WRIST),
       ("hand_base_link", (0.0, 0.035, 0.0), 0.03),
       ("hand_middle_finger_lower_link", (0.025, 0.0, 0.0), R_FINGER),
       ("hand_middle_finger_upper_link", (0.015, 0.0, 0.0), R_FINGER),
       ("hand_left_crank", (0.0, 0.0, 0.0), 0.01),
       ("hand_left_coupler", (0.008, 0.01, 0.03), R_FINGER),
       ("hand_left_finger_lower_link", (0.025, 0.0, 0.0), R_FINGER),
       ("hand_left_finger_upper_link", (0.025, 0.0, 0.0), R_FINGER),
       ("hand_right_crank", (0.0, 0.0, 0.0), 0.01),
       ("hand_right_coupler", (-0.008, 0.01, 0.03), R_FINGER),
       ("hand_right_finger_lower_link", (0.025, 0.0, 0.0), R_FINGER),
       ("hand_right_finger_upper_link", (0.025, 0.0, 0.0), R_FINGER),
       ("hand_cableCanal_link", (0.0, 0.0, 0.0), 0.03)]
)


def spheres_for(robot: str = "robotino"):
    if robot != "robotino":
        raise KeyError(robot)
    return [(n, tuple(float(v) for v in c), float(r)) for n, c, r in ROBOTINO_SPHERES]
