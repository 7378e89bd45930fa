#!/usr/bin/env python
"""Write the compiled robot description (what sqmp_create takes) as a text file the C++ shim loads.

    python tools/export_robot_desc.py [out_path]

Input: squirrel_motion_planner_b200/data/robotino_plan.json (link table parsed from the reference URDF/SRDF by
tools/gen_robot_fixture.py + the authored sphere model).  Output: data/robotino_plan.sqmpdesc, format "sqmpdesc 2".
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from squirrel_motion_planner_b200 import robot_desc as rd  # noqa: E402


def fmt(v):
    return " ".join("%.17g" % float(x) for x in v)


def main():
    lt, fx = rd.load_fixture()
    spheres = [(n, tuple(c), r) for n, c, r in fx["spheres"]]
    cr = rd.compile_robot(lt, spheres)
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(rd.DEFAULT_FIXTURE), "robotino_plan.sqmpdesc")
    L = ["sqmpdesc 2", "n_seg %d" % cr.n_seg]
    for s in range(cr.n_seg):
        L.append("seg %d %d %s %s" % (cr.seg_type[s], cr.seg_qidx[s], fmt(cr.seg_pre[s]), fmt(cr.seg_axis[s])))
    L.append("root_offset " + fmt(cr.root_offset))
    L.append("ee_tip " + fmt(cr.ee_tip))
    L.append("n_links %d" % lt.n)
    for i, n in enumerate(lt.names):
        L.append("link %d %s" % (i, n))
    L.append("octomap_disabled %d %s" % (len(lt.octomap_disabled), " ".join(lt.octomap_disabled)))
    L.append("n_spheres %d" % cr.n_spheres)
    for i in range(cr.n_spheres):
        L.append("sphere %d %d %s %.17g" % (cr.sph_seg[i], cr.sph_link[i], fmt(cr.sph_center[i]), cr.sph_radius[i]))
    rows = [(i, int(v)) for i, v in enumerate(cr.link_pair_mask) if int(v)]
    L.append("n_link_pair_rows %d" % len(rows))
    for i, v in rows:
        L.append("link_pairs %d %x" % (i, v))
    L.append("self_const_collision %d" % cr.self_const_collision)
    with open(out, "w") as f:
        f.write("\n".join(L) + "\n")
    print("wrote", out, "(%d segments, %d spheres, %d link-pair rows)" % (cr.n_seg, cr.n_spheres, len(rows)))


if __name__ == "__main__":
    main()
