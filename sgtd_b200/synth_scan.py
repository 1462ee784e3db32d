"""Synthetic SemanticKITTI-shaped labelled scans.

A 64-beam spinning LiDAR is ray-cast against a procedural street scene: a ground
plane carrying road / sidewalk / terrain / other-ground regions, vertical
cylinders (poles, trunks), vertical rectangles (building facades, fences,
traffic signs) and boxes (cars).  Output has the layout of a KITTI `.bin`
(float32 x,y,z,intensity) and `.label` (uint32, lo16 = train id 0..19, hi16 =
instance id, 0 here: forces the DCVC branch like SPVNAS predictions,
R/src/get_json.cpp:160-226).  Data generation only; not part of the hot path.
"""
import numpy as np

SENSOR_H = 1.73
# train ids (R/src/get_json.cpp:13-33)
CAR, ROAD, SIDEWALK, OTHER_GROUND, BUILDING, FENCE, VEGETATION, TRUNK, TERRAIN, POLE, SIGN = 0, 8, 10, 11, 12, 13, 14, 15, 16, 17, 18


def make_scene(seed, extent=90.0, n_poles=30, n_trunks=30, n_signs=12, n_facades=14, n_fences=8, n_cars=10):
    """Primitives in the sensor frame (sensor at origin, ground at z = -SENSOR_H)."""
    rng = np.random.default_rng(seed)

    def ring(n, rmin, rmax):
        r = rng.uniform(rmin, rmax, n)
        a = rng.uniform(-np.pi, np.pi, n)
        return np.column_stack([r * np.cos(a), r * np.sin(a)])

    cyl = []  # (x, y, radius, z0, z1, label)
    for p in ring(n_poles, 6, extent):
        cyl.append((p[0], p[1], rng.uniform(0.08, 0.18), -SENSOR_H, rng.uniform(2.5, 6.0) - SENSOR_H, POLE))
    for p in ring(n_trunks, 6, extent):
        cyl.append((p[0], p[1], rng.uniform(0.15, 0.35), -SENSOR_H, rng.uniform(1.5, 3.0) - SENSOR_H, TRUNK))
    rect = []  # vertical rectangles: (x0, y0, x1, y1, z0, z1, label)
    for p in ring(n_facades, 12, extent):
        ang = rng.uniform(0, np.pi)
        L = rng.uniform(6, 30)
        d = np.array([np.cos(ang), np.sin(ang)]) * L / 2
        rect.append((p[0] - d[0], p[1] - d[1], p[0] + d[0], p[1] + d[1], -SENSOR_H, rng.uniform(4, 12) - SENSOR_H, BUILDING))
    for p in ring(n_fences, 8, extent * 0.7):
        ang = rng.uniform(0, np.pi)
        L = rng.uniform(5, 20)
        d = np.array([np.cos(ang), np.sin(ang)]) * L / 2
        rect.append((p[0] - d[0], p[1] - d[1], p[0] + d[0], p[1] + d[1], -SENSOR_H, 1.3 - SENSOR_H, FENCE))
    for p in ring(n_signs, 5, extent * 0.6):
        ang = rng.uniform(0, np.pi)
        d = np.array([np.cos(ang), np.sin(ang)]) * 0.4
        z0 = rng.uniform(1.8, 2.6) - SENSOR_H
        rect.append((p[0] - d[0], p[1] - d[1], p[0] + d[0], p[1] + d[1], z0, z0 + 0.8, SIGN))
    for p in ring(n_cars, 4, extent * 0.5):  # cars as two crossing rectangles (enough to occlude)
        ang = rng.uniform(0, np.pi)
        for a2, L in ((ang, 4.2), (ang + np.pi / 2, 1.8)):
            d = np.array([np.cos(a2), np.sin(a2)]) * L / 2
            rect.append((p[0] - d[0], p[1] - d[1], p[0] + d[0], p[1] + d[1], -SENSOR_H, 1.5 - SENSOR_H, CAR))
    patches = np.column_stack([ring(40, 5, extent), rng.uniform(3, 12, 40), rng.choice([TERRAIN, OTHER_GROUND, VEGETATION], 40, p=[0.5, 0.2, 0.3])])
    return dict(cyl=np.array(cyl), rect=np.array(rect), patches=patches, road_half=rng.uniform(3.0, 4.5),
                walk=rng.uniform(1.5, 2.5))


def render(scene, seed, n_beams=64, n_az=1875, max_range=120.0, noise=0.02):
    """Ray-cast the scene.  Returns (points float32 [N,4], labels uint32 [N]) in firing order
    (azimuth-major, beam-minor), dropping rays without a return."""
    rng = np.random.default_rng(seed)
    pitch = np.radians(np.linspace(2.0, -24.8, n_beams))
    az = np.linspace(-np.pi, np.pi, n_az, endpoint=False) + rng.uniform(0, 2 * np.pi / n_az)
    A, P = np.meshgrid(az, pitch, indexing="ij")
    A, P = A.ravel(), P.ravel()
    dx, dy, dz = np.cos(P) * np.cos(A), np.cos(P) * np.sin(A), np.sin(P)
    best = np.full(A.shape, np.inf)
    lab = np.full(A.shape, 255, np.int64)
    # ground
    with np.errstate(divide="ignore", invalid="ignore"):
        tg = np.where(dz < -1e-6, -SENSOR_H / dz, np.inf)
    gx, gy = dx * tg, dy * tg
    glab = np.full(A.shape, TERRAIN, np.int64)
    glab[np.abs(gy) < scene["road_half"]] = ROAD
    band = (np.abs(gy) >= scene["road_half"]) & (np.abs(gy) < scene["road_half"] + scene["walk"])
    glab[band] = SIDEWALK
    off = ~(band | (np.abs(gy) < scene["road_half"]))
    for (px, py, pr, pl) in scene["patches"]:
        m = off & ((gx - px) ** 2 + (gy - py) ** 2 < pr * pr)
        glab[m] = int(pl)
    hit = tg < best
    best[hit], lab[hit] = tg[hit], glab[hit]
    # cylinders (2-D circle intersection, then height check)
    dxy2 = dx * dx + dy * dy
    for (cx, cy, cr, z0, z1, cl) in scene["cyl"]:
        b = dx * cx + dy * cy
        c = cx * cx + cy * cy - cr * cr
        disc = b * b - dxy2 * c
        ok = disc > 0
        t = np.where(ok, (b - np.sqrt(np.where(ok, disc, 0))) / dxy2, np.inf)
        z = dz * t
        ok &= (t > 0.5) & (z >= z0) & (z <= z1) & (t < best)
        best[ok], lab[ok] = t[ok], int(cl)
    # vertical rectangles
    for (x0, y0, x1, y1, z0, z1, cl) in scene["rect"]:
        ex, ey = x1 - x0, y1 - y0
        den = dx * ey - dy * ex
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (x0 * ey - y0 * ex) / den
            u = (x0 * dy - y0 * dx) / den
        z = dz * t
        ok = (np.abs(den) > 1e-9) & (t > 0.5) & (u >= 0) & (u <= 1) & (z >= z0) & (z <= z1) & (t < best)
        best[ok], lab[ok] = t[ok], int(cl)
    keep = np.isfinite(best) & (best < max_range) & (lab != 255)
    r = best[keep] + rng.normal(0, noise, keep.sum())
    pts = np.column_stack([dx[keep] * r, dy[keep] * r, dz[keep] * r, rng.uniform(0, 1, keep.sum())]).astype(np.float32)
    return pts, lab[keep].astype(np.uint32)


def make_scan(seed, **kw):
    scene = make_scene(seed, **{k: v for k, v in kw.items() if k.startswith("n_") and k not in ("n_beams", "n_az")} or {})
    return render(scene, seed + 1, **{k: v for k, v in kw.items() if k in ("n_beams", "n_az", "noise", "max_range")})
