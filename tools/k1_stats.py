"""Development aid: statistics of the K1 filter cascade on the config-2 workload (numpy, CPU).
For every (state, sphere): cell box-to-box distance (what the byte distance field stores), exact distance, verdict."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy import ndimage
from squirrel_motion_planner_b200 import robot_desc as rd, synthetic as syn


def fk_spheres(cr, q):
    """world sphere centres [S, n, 3] by the compiled chain (vectorised numpy)."""
    S = len(q)
    R = np.tile(np.eye(3), (S, 1, 1)); p = np.tile(cr.root_offset, (S, 1))
    out = np.empty((S, cr.n_spheres, 3))
    def emit(seg):
        idx = np.nonzero(cr.sph_seg == seg)[0]
        for i in idx:
            out[:, i] = np.einsum('sij,j->si', R, cr.sph_center[i]) + p
    emit(-1)
    for k in range(cr.n_seg):
        pre = cr.seg_pre[k]; Rp = pre[:9].reshape(3, 3); tp = pre[9:]
        p = p + np.einsum('sij,j->si', R, tp); R = R @ Rp
        qv = q[:, cr.seg_qidx[k]]; ax = cr.seg_axis[k]
        if cr.seg_type[k] == rd.JOINT_TRANS:
            p = p + np.einsum('sij,j->si', R, ax) * qv[:, None]
        else:
            c, s = np.cos(qv), np.sin(qv)
            K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
            M = np.eye(3)[None] + s[:, None, None] * K[None] + (1 - c)[:, None, None] * (K @ K)[None]
            R = R @ M
        emit(k)
    return out


def main():
    S = int(os.environ.get("S", "20000"))
    lt, fx = rd.load_fixture()
    spheres = [(n, tuple(c), r) for n, c, r in fx["spheres"]]
    cr = rd.compile_robot(lt, spheres)
    w = syn.box_world(g=256, n_boxes=64, seed=1234)
    origin = w.origin + np.array([0, 0, rd.OCTREE_Z_OFFSET])
    occ = w.occupancy()  # [z,y,x]
    q = syn.state_batch(1 << 20, seed=42)[:: (1 << 20) // S][:S]
    C = fk_spheres(cr, q)
    U = (C - origin) / w.res            # grid units
    rg = cr.sph_radius / w.res
    print("spheres", cr.n_spheres, "segs", cr.n_seg, "radii(cells)", np.unique(np.round(rg, 2)))
    print("spheres per seg", [(s, int((cr.sph_seg == s).sum())) for s in range(-1, cr.n_seg)])
    # exact point-to-occupied distance via brute force around each centre (cap 7 cells)
    gz, gy, gx = occ.shape
    I = np.floor(U).astype(int)
    inside = np.all((I >= 0) & (I < [gx, gy, gz]), axis=2)
    print("inside frac", inside.mean())
    # cell-level box-box distance: EDT on cells gives centre-centre dist; box-box = sqrt(sum max(|d|-1,0)^2): compute by brute force window
    W = 7
    pad = np.pad(occ, W, constant_values=False)
    offs = np.arange(-W, W + 1)
    dz, dy, dx = np.meshgrid(offs, offs, offs, indexing='ij')
    bb2 = (np.maximum(np.abs(dx) - 1, 0) ** 2 + np.maximum(np.abs(dy) - 1, 0) ** 2 + np.maximum(np.abs(dz) - 1, 0) ** 2).astype(float)
    n_amb = n_hit = n_tot = 0
    per_sphere = np.zeros((cr.n_spheres, 4))
    state_hit = np.zeros(S, bool)
    rows_scanned = 0
    for s in range(S):
        for i in range(cr.n_spheres):
            if not inside[s, i]:
                continue
            ix, iy, iz = I[s, i]
            win = pad[iz:iz + 2 * W + 1, iy:iy + 2 * W + 1, ix:ix + 2 * W + 1]
            if not win.any():
                per_sphere[i, 0] += 1
                continue
            bbd = np.sqrt(bb2[win].min())
            e = np.floor(4 * bbd)
            thr = np.floor(4 * rg[i])
            per_sphere[i, 0] += 1
            if e > thr:
                continue
            # exact
            f = U[s, i] - I[s, i]
            ex = np.maximum(np.maximum(dx - f[0], f[0] - (dx + 1)), 0)
            ey = np.maximum(np.maximum(dy - f[1], f[1] - (dy + 1)), 0)
            ez = np.maximum(np.maximum(dz - f[2], f[2] - (dz + 1)), 0)
            d = np.sqrt((ex ** 2 + ey ** 2 + ez ** 2)[win].min())
            per_sphere[i, 1] += 1
            nr = int(np.ceil(rg[i])) * 2 + 2
            per_sphere[i, 3] += nr * nr
            if d <= rg[i]:
                per_sphere[i, 2] += 1
                state_hit[s] = True
    np.set_printoptions(linewidth=200, suppress=True)
    print("state hit frac", state_hit.mean())
    print("idx seg r(cells)  tested  ambiguous(frac)  hit(frac of amb)  rows/amb")
    for i in range(cr.n_spheres):
        t, a, h, r = per_sphere[i]
        print("%3d %3d %5.2f %8d  %.3f  %.3f  %5.0f   %s" % (i, cr.sph_seg[i], rg[i], t, a / max(t, 1), h / max(a, 1), r / max(a, 1), lt.names[cr.sph_link[i]]))
    tot = per_sphere.sum(0)
    print("total tested %d, ambiguous %d (%.3f per state), hits %d, rows per state %.1f" % (tot[0], tot[1], tot[1] / S, tot[2], tot[3] / S))


if __name__ == "__main__":
    main()
