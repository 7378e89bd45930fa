#!/usr/bin/env python
"""Generate squirrel_motion_planner_b200/data/robotino_plan.json from the reference's URDF/SRDF.

Run in the build container (needs /root/reference); the result is committed so that the GPU box,
which has no /root/reference, can build the same robot description.  Nothing but kinematic
constants, the SRDF pair table and a few published poses is taken over — no reference source code.

    python tools/gen_robot_fixture.py [--reference /root/reference]
"""
import argparse
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from squirrel_motion_planner_b200 import robot_desc, sphere_model  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=robot_desc.DEFAULT_FIXTURE)
    ap.add_argument("--refit", action="store_true",
                    help="re-run the covering sphere fit (sphere_model.build_model, ~1 min) and refreeze data/robotino_spheres.json")
    a = ap.parse_args()
    cfg = os.path.join(a.reference, "squirrel_8dof_planner", "config")
    urdf = open(os.path.join(cfg, "robotino_plan.urdf")).read()
    srdf = open(os.path.join(cfg, "robotino_plan.srdf")).read()
    lt = robot_desc.link_table_from_urdf(urdf, srdf)
    rigid = robot_desc.disable_rigid_pairs(lt)
    print("rigid pairs taken out of the self-collision matrix:", len(rigid))
    if a.refit or not os.path.exists(sphere_model.FROZEN):
        sph = sphere_model.build_model(lt, progress=print)
        with open(sphere_model.FROZEN, "w") as f:
            json.dump({"generator": "tools/gen_robot_fixture.py --refit (sphere_model.build_model)",
                       "inflation_m": {"primitive": sphere_model.P_PRIMITIVE, "bottom": sphere_model.P_BOTTOM,
                                       "arm": sphere_model.P_ARM, "hand": sphere_model.P_HAND},
                       "spheres": [[n, list(c), r] for n, c, r in sph]}, f, indent=0)
        print("froze", sphere_model.FROZEN, len(sph), "spheres")

    # published poses (fixtures for sanity tests): folding trajectory and folded arm
    fold = open(os.path.join(cfg, "folding_poses_tuw-robotino2.yaml")).read()
    nums = [float(t) for t in re.findall(r"-?\d+\.\d+(?:[eE]-?\d+)?", fold.split(":", 1)[1])]
    folding = [nums[i:i + 5] for i in range(0, len(nums) - len(nums) % 5, 5)]
    params = open(os.path.join(cfg, "parameters.yaml")).read()
    folded = [float(t) for t in re.search(r"pose_folded_arm:\s*\[([^\]]*)\]", params).group(1).split(",")]

    out = {
        "source": {
            "urdf": "squirrel_8dof_planner/config/robotino_plan.urdf",
            "srdf": "squirrel_8dof_planner/config/robotino_plan.srdf",
            "folding": "squirrel_8dof_planner/config/folding_poses_tuw-robotino2.yaml:1-24",
            "folded": "squirrel_8dof_planner/config/parameters.yaml:37",
            "generator": "tools/gen_robot_fixture.py",
        },
        "link_table": lt.to_json(),
        "spheres": [[n, list(c), r] for n, c, r in sphere_model.spheres_for("robotino")],
        "folding_poses": folding,
        "pose_folded_arm": folded,
        "planner_params": {  # planner_parameters_robotino.yaml:3-15
            "near_threshold_interpolation": 1.5, "num_traj_segments_interp": 20, "max_near_nodes": 20,
            "unconstraint_extend_step_factor": 0.6},
    }
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", a.out, "links", lt.n, "shaped", int(lt.shaped.sum()), "disabled pairs", len(lt.disabled_pairs),
          "folding poses", len(folding))


if __name__ == "__main__":
    main()
