"""grade the authored collision-sphere model (squirrel_motion_planner_b200/sphere_model.py) against the
reference's link shapes and print, for the six URDF primitives, a covering fit of the same radius next to it.
usage: python tools/sphere_model_report.py [--radius R]   (R: also print covering fits of every primitive at radius R)"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from squirrel_motion_planner_b200 import robot_desc as rd, sphere_fit as sf  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--radius", type=float, default=0.0)
    a = ap.parse_args()
    lt, fx = rd.load_fixture()
    spheres = [(n, tuple(c), r) for n, c, r in fx["spheres"]]
    print("%-34s %-28s %3s  %-22s %-22s" % ("link", "shape", "n", "authored cov / protr", "covering fit n / protr"))
    for row in sf.report(lt, spheres):
        au = row.get("authored")
        ft = row.get("fit")
        print("%-34s %-28s %3d  %-22s %-22s" % (
            row["link"], row["shape"], row["n_spheres"],
            "%.3f / %.3f m" % (au["covered"], au["protrusion"]) if au else "-",
            "%d / %.3f m" % (ft["n_spheres"], ft["protrusion"]) if ft else "-"))
    if a.radius > 0.0:
        for i, name in enumerate(lt.names):
            shape = sf.parse_shape(lt.shape_desc[i]) if lt.shaped[i] else None
            if shape is None or min(shape[1]) <= 0.005:
                continue
            c = sf.fit_primitive(shape, a.radius)
            print("# %s: %d spheres of r=%.3f cover it (%s)" % (name, len(c), a.radius, sf.grade(shape, c, np.full(len(c), a.radius))))
            for v in c:
                print('    ("%s", (%.4f, %.4f, %.4f), %.3f),' % (name, v[0], v[1], v[2], a.radius))


if __name__ == "__main__":
    main()
