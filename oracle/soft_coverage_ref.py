"""NumPy restatement of ``AuxiliaryCameraRewards.compute_soft_coverage_scores``
(mate/wrappers/auxiliary_camera_rewards.py:182-233) on top of the OUTER field-of-view boundary
(``Camera.add_obstacles`` / ``boundary_between(outer=True)``, mate/entities.py:362-479, 484-511) -- TEST
INFRASTRUCTURE: the checker of the CUDA kernel, itself pinned against the reference's values in
``tests/golden/aux_*.npz`` (``a_out_soft_matrix``)."""

import numpy as np


def norm_angle(a):
    return (a + 180.0) % 360.0 - 180.0


def outer_samples(cam, obs_xyr, rmax):
    """(angle, norm) of every sample ray of the outer boundary, before the obstacles cut them."""
    angles = list(np.linspace(-180.0, 180.0, num=360, endpoint=False))
    norms = [rmax] * 360
    discs = []
    for x, y, r in obs_xyr:
        rel = np.array([x, y]) - cam
        d = float(np.hypot(rel[0], rel[1]))
        if not d < rmax + r:      # entities.py:363-368
            continue
        discs.append((rel, d, r))
    for rel, d, r in discs:
        if r > d:                 # camera inside a disc (entities.py:378-388): the boundary collapses to the camera
            return np.array([-180.0, -90.0, 0.0, 90.0]), np.zeros(4), discs
        ang = np.rad2deg(np.arctan2(rel[1], rel[0]))
        half = np.rad2deg(np.arcsin(r / d))
        max_rho = min(rmax, d + r)
        left, right = ang - half, ang + half
        for a in np.linspace(left, right, num=max(16, int(2 * half)) + 1, endpoint=True):   # entities.py:419-429
            angles.append(norm_angle(float(a)))
            norms.append(max_rho)
        near_rho, far_rho = min(rmax, np.sqrt(d * d + r * r)), rmax                            # entities.py:431-448
        for edge, far_angle in ((left, left - 0.01), (right, right + 0.01)):
            na, fa = np.deg2rad(norm_angle(edge)), np.deg2rad(norm_angle(far_angle))
            near = near_rho * np.array([np.cos(na), np.sin(na)])
            far = far_rho * np.array([np.cos(fa), np.sin(fa)])
            for t in np.linspace(0.0, 1.0, num=21, endpoint=True):
                v = (1.0 - t) * near + t * far
                angles.append(float(np.rad2deg(np.arctan2(v[1], v[0]))))
                norms.append(float(np.hypot(v[0], v[1])))
    return np.array(angles), np.array(norms), discs


def obstruct_outer(angles, norms, discs, tangent_eps=1e-9):
    """Obstacle.obstruct(ray, outer=True) (entities.py:158-184) for all rays against all discs: a ray that crosses
    a disc ends at the FAR side of it."""
    rad = np.deg2rad(angles)
    ux, uy = np.cos(rad), np.sin(rad)
    out = norms.copy()
    for rel, d, r in discs:
        proj = rel[0] * ux + rel[1] * uy
        reach = ~(d >= out + r) & (out > 0.0)
        cos = np.minimum(1.0, proj / d)
        perp = d * np.sqrt(np.maximum(1.0 - cos * cos, 0.0))
        # tangent_eps > 0: exactly tangent rays are not cut (exact arithmetic, the convention of the CUDA path);
        # tangent_eps < 0: they are (the reference decides them by rounding noise, DESIGN.md "Tangent rays")
        hit = reach & (proj >= 0.0) & (r > perp * (1.0 + tangent_eps))
        new = np.maximum(0.0, d * cos + np.sqrt(np.maximum(r * r - perp * perp, 0.0)))
        out = np.where(hit & (new < out), new, out)
    return out


def soft_coverage_matrix(cam_xy, cam_phi, cam_theta, rmax, area_product, obs_xyr, tgt_xy, mask_ct, inner_tables, tangent_eps=1e-9):
    """[Nc, Nt] matrix of the reference.  inner_tables[c] = (phi, rho) of the camera's inner polyline
    (``sight_range_func.x/.y``), used for the two sector edges like ``boundary_between`` does."""
    nc, nt = len(cam_xy), len(tgt_xy)
    out = np.zeros((nc, nt))
    for c in range(nc):
        theta, phi = float(cam_theta[c]), float(cam_phi[c])
        sight_range = np.sqrt(area_product / theta)
        dist_max = sight_range / (1.0 + 1.0 / np.sin(np.deg2rad(theta / 2.0))) if theta < 180.0 else sight_range / 2.0
        angles, norms, discs = outer_samples(np.asarray(cam_xy[c], dtype=np.float64), obs_xyr, rmax)
        rhos_all = obstruct_outer(angles, norms, discs, tangent_eps)
        left = norm_angle(phi - theta / 2.0)
        right = left + theta
        if right <= 180.0:
            inside = (left < angles) & (angles < right)
        else:   # the polyline's closing sample (phi0 + 360 = +180) stands for the -180 ray
            inside = (angles > left) | (angles < right - 360.0) | (angles == -180.0)
        tphi, trho = inner_tables[c]
        rho_left = float(np.interp(norm_angle(left), tphi, trho))
        rho_right = float(np.interp(norm_angle(right), tphi, trho))
        phis = np.concatenate([[left] * 16, [left], angles[inside], [right], [right] * 16])
        rhos = np.concatenate([np.linspace(0.0, rho_left, num=16, endpoint=False), [rho_left], rhos_all[inside], [rho_right],
                               np.linspace(0.0, rho_right, num=16, endpoint=False)])
        xs, ys = rhos * np.cos(np.deg2rad(phis)), rhos * np.sin(np.deg2rad(phis))
        for t in range(nt):
            direction = np.asarray(tgt_xy[t], dtype=np.float64) - np.asarray(cam_xy[c], dtype=np.float64)
            dist = np.hypot(direction[0] - xs, direction[1] - ys).min()
            out[c, t] = (dist if mask_ct[c, t] else -dist) / dist_max
    return out


def soft_coverage_explained(cam_xy, cam_phi, cam_theta, rmax, area_product, obs_xyr, tgt_xy, mask_ct, inner_tables, gold, tol=1e-6):
    """Is every entry of the reference's matrix `gold` reachable by SOME outcome of the exactly tangent rays?

    The reference cuts a boundary ray that is exactly tangent to an obstacle or not depending on the last bit of NumPy's
    sin / cos (DESIGN.md "Tangent rays"), independently per ray.  A ray whose end point differs between "all tangent
    rays cut" and "none cut" is such a ray; it contributes one of two candidate points.  An entry min_k |target - p_k| is
    reachable iff (i) it equals the distance to a point some outcome contains and (ii) every ambiguous ray has an option
    that is not closer than it (and no unambiguous point is closer).  Returns ([Nc, Nt] bool reachable, [Nc, Nt] bool
    equal to the none-cut convention of the CUDA path)."""
    nc, nt = len(cam_xy), len(tgt_xy)
    ok = np.zeros((nc, nt), dtype=bool)
    same = np.zeros((nc, nt), dtype=bool)
    for c in range(nc):
        theta, phi = float(cam_theta[c]), float(cam_phi[c])
        sight_range = np.sqrt(area_product / theta)
        dist_max = sight_range / (1.0 + 1.0 / np.sin(np.deg2rad(theta / 2.0))) if theta < 180.0 else sight_range / 2.0
        cam = np.asarray(cam_xy[c], dtype=np.float64)
        angles, norms, discs = outer_samples(cam, obs_xyr, rmax)
        rho0 = obstruct_outer(angles, norms, discs, +1e-9)    # no tangent ray cut
        rho1 = obstruct_outer(angles, norms, discs, -1e-9)    # every tangent ray cut
        left = norm_angle(phi - theta / 2.0)
        right = left + theta
        if right <= 180.0:
            inside = (left < angles) & (angles < right)
        else:
            inside = (angles > left) | (angles < right - 360.0) | (angles == -180.0)
        tphi, trho = inner_tables[c]
        rho_left = float(np.interp(norm_angle(left), tphi, trho))
        rho_right = float(np.interp(norm_angle(right), tphi, trho))
        edge_phi = np.concatenate([[left] * 17, [right] * 17])
        edge_rho = np.concatenate([np.linspace(0.0, rho_left, num=16, endpoint=False), [rho_left], [rho_right],
                                   np.linspace(0.0, rho_right, num=16, endpoint=False)])
        a_in = np.deg2rad(angles[inside])
        amb = np.abs(rho0[inside] - rho1[inside]) > 1e-9
        fixed_x = np.concatenate([edge_rho * np.cos(np.deg2rad(edge_phi)), (rho0[inside] * np.cos(a_in))[~amb]])
        fixed_y = np.concatenate([edge_rho * np.sin(np.deg2rad(edge_phi)), (rho0[inside] * np.sin(a_in))[~amb]])
        ax0, ay0 = (rho0[inside] * np.cos(a_in))[amb], (rho0[inside] * np.sin(a_in))[amb]
        ax1, ay1 = (rho1[inside] * np.cos(a_in))[amb], (rho1[inside] * np.sin(a_in))[amb]
        for t in range(nt):
            d = np.asarray(tgt_xy[t], dtype=np.float64) - cam
            d_fixed = np.hypot(d[0] - fixed_x, d[1] - fixed_y).min()
            d0, d1 = np.hypot(d[0] - ax0, d[1] - ay0), np.hypot(d[0] - ax1, d[1] - ay1)
            want = abs(float(gold[c, t])) * dist_max
            sign_ok = (gold[c, t] >= 0) == bool(mask_ct[c, t]) or want < tol
            upper = min(d_fixed, np.maximum(d0, d1).min() if len(d0) else np.inf)
            attained = abs(want - d_fixed) < tol * dist_max or (len(d0) and (np.abs(d0 - want).min() < tol * dist_max or np.abs(d1 - want).min() < tol * dist_max))
            ok[c, t] = sign_ok and want <= upper + tol * dist_max and bool(attained)
            none_cut = min(d_fixed, d0.min() if len(d0) else np.inf)
            same[c, t] = abs(want - none_cut) < tol * dist_max
    return ok, same


def after_step_cameras(g, i):
    """Camera.simulate (entities.py:347-360) on the recorded state / action of sample i of an aux fixture."""
    cam = g['cfg_camera']
    act = g['a_cam_act'][i]
    phi = norm_angle(g['a_cam_phi'][i] + np.clip(act[:, 0], -cam[3], cam[3]))
    theta = np.clip(g['a_cam_theta'][i] + np.clip(act[:, 1], -cam[4], cam[4]), cam[1], 180.0)
    return phi, theta
