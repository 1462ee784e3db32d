"""Synthetic labelled-scan SEQUENCES (BASELINE.json configs[1]/[4]-shaped workloads).

A persistent street world (Manhattan grid from synth.make_world, street furniture along
the grid lines) is ray-cast from every trajectory pose with a 64-beam spinning LiDAR
model, so consecutive and revisiting scans observe the same landmarks from different
viewpoints -- what stage 1 (instance extraction) needs to be exercised in situ.
The ray-caster is written in torch so the bench can generate thousands of scans on the
GPU in minutes; it runs on the CPU too (tests).  Data generation only -- not part of the
hot path, never imported by the library.
"""
import numpy as np
import torch

from . import synth
from .synth_scan import (BUILDING, CAR, FENCE, OTHER_GROUND, POLE, ROAD, SENSOR_H, SIDEWALK, SIGN, TERRAIN, TRUNK,
                         VEGETATION)


def make_street_world(n_scans, seed, spacing=0.8, block=120.0):
    base = synth.make_world(n_scans, seed, spacing=spacing, block=block, density=1e-6)
    rng = np.random.default_rng(seed + 77)
    G = base["G"]
    cyl, rect = [], []

    def add_along(p0, p1):
        d = p1 - p0
        L = float(np.hypot(*d))
        u = d / L
        nrm = np.array([-u[1], u[0]])
        for side in (-1.0, 1.0):
            s = rng.uniform(5, 20)
            while s < L - 5:                      # poles / trunks / signs near the kerb
                kind = rng.choice(3, p=[0.45, 0.4, 0.15])
                c = p0 + u * s + nrm * side * rng.uniform(5.5, 8.5)
                if kind == 0:
                    cyl.append((c[0], c[1], rng.uniform(0.08, 0.16), 0.0, rng.uniform(3.0, 6.5), POLE))
                elif kind == 1:
                    cyl.append((c[0], c[1], rng.uniform(0.18, 0.35), 0.0, rng.uniform(1.8, 3.2), TRUNK))
                else:
                    e = u * 0.45
                    z0 = rng.uniform(1.9, 2.6)
                    rect.append((c[0] - e[0], c[1] - e[1], c[0] + e[0], c[1] + e[1], z0, z0 + 0.8, SIGN))
                s += rng.uniform(9, 28)
            s = rng.uniform(8, 14)
            while s < L - 8:                      # facades / fences set back from the street
                ln = rng.uniform(8, 28)
                if s + ln > L - 6:
                    break
                setback = rng.uniform(10, 17)
                a = p0 + u * s + nrm * side * setback
                b = a + u * ln
                if rng.random() < 0.75:
                    rect.append((a[0], a[1], b[0], b[1], 0.0, rng.uniform(5, 12), BUILDING))
                else:
                    rect.append((a[0], a[1], b[0], b[1], 0.0, 1.3, FENCE))
                s += ln + rng.uniform(3, 14)
            s = rng.uniform(10, 40)
            while s < L - 10:                     # parked cars (occluders, class skipped by the pipeline)
                c = p0 + u * s + nrm * side * 3.3
                e = u * 2.1
                rect.append((c[0] - e[0], c[1] - e[1], c[0] + e[0], c[1] + e[1], 0.0, 1.5, CAR))
                s += rng.uniform(12, 60)

    for i in range(G):
        for j in range(G - 1):
            add_along(np.array([i * block, j * block], float), np.array([i * block, (j + 1) * block], float))
            add_along(np.array([j * block, i * block], float), np.array([(j + 1) * block, i * block], float))
    npatch = 6 * G * G
    patches = np.column_stack([rng.uniform(0, (G - 1) * block, (npatch, 2)), rng.uniform(3, 10, npatch),
                               rng.choice([OTHER_GROUND, VEGETATION], npatch, p=[0.5, 0.5])])
    return dict(poses=base["poses"], G=G, block=block, cyl=np.array(cyl, np.float64), rect=np.array(rect, np.float64),
                patches=patches, road_half=3.5, walk=2.0, seed=seed)


def render_at(world, pose, seed, device="cpu", n_beams=64, n_az=1875, max_range=100.0, noise=0.02, reach=105.0):
    """One scan from `pose` = (x, y, yaw).  Returns (points float32 [N,4] tensor, labels int64 [N] tensor)
    on `device`, in firing order (azimuth-major)."""
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    x0, y0, yaw = float(pose[0]), float(pose[1]), float(pose[2])
    c, s = np.cos(yaw), np.sin(yaw)

    def to_sensor(xy):
        d = xy - np.array([x0, y0])
        return np.column_stack([c * d[:, 0] + s * d[:, 1], -s * d[:, 0] + c * d[:, 1]])

    dt = torch.float64
    pitch = torch.deg2rad(torch.linspace(2.0, -24.8, n_beams, dtype=dt))
    az = torch.linspace(-np.pi, np.pi, n_az + 1, dtype=dt)[:-1] + float(torch.rand(1, generator=g)) * 2 * np.pi / n_az
    A = az.repeat_interleave(n_beams).to(device)
    P = pitch.repeat(n_az).to(device)
    dx, dy, dz = torch.cos(P) * torch.cos(A), torch.cos(P) * torch.sin(A), torch.sin(P)
    inf = torch.tensor(float("inf"), dtype=dt, device=device)
    # ---- ground (labels from world coordinates of the hit) ----
    tg = torch.where(dz < -1e-6, -SENSOR_H / dz, inf)
    gx, gy = dx * tg, dy * tg
    wx, wy = x0 + c * gx - s * gy, y0 + s * gx + c * gy
    B = world["block"]
    dline = torch.minimum((wx - torch.round(wx / B) * B).abs(), (wy - torch.round(wy / B) * B).abs())
    glab = torch.full_like(A, TERRAIN, dtype=torch.int64)
    off_road = dline >= world["road_half"] + world["walk"]
    pc = world["patches"]
    near = pc[np.hypot(pc[:, 0] - x0, pc[:, 1] - y0) < reach + 12]
    for (px, py, pr, pl) in near:
        glab = torch.where(off_road & ((wx - px) ** 2 + (wy - py) ** 2 < pr * pr), torch.tensor(int(pl), device=device), glab)
    glab = torch.where(dline < world["road_half"] + world["walk"], torch.tensor(SIDEWALK, device=device), glab)
    glab = torch.where(dline < world["road_half"], torch.tensor(ROAD, device=device), glab)
    best, lab = tg, glab
    # ---- cylinders ----
    cy = world["cyl"]
    cy = cy[np.hypot(cy[:, 0] - x0, cy[:, 1] - y0) < reach]
    if len(cy):
        cs = to_sensor(cy[:, :2])
        cx_, cy_ = torch.tensor(cs[:, 0], device=device), torch.tensor(cs[:, 1], device=device)
        cr = torch.tensor(cy[:, 2], device=device)
        z0 = torch.tensor(cy[:, 3] - SENSOR_H, device=device)
        z1 = torch.tensor(cy[:, 4] - SENSOR_H, device=device)
        cl = torch.tensor(cy[:, 5].astype(np.int64), device=device)
        dxy2 = (dx * dx + dy * dy)[:, None]
        for k0 in range(0, len(cy), 64):
            sl = slice(k0, k0 + 64)
            b = dx[:, None] * cx_[None, sl] + dy[:, None] * cy_[None, sl]
            cc = (cx_[sl] ** 2 + cy_[sl] ** 2 - cr[sl] ** 2)[None, :]
            disc = b * b - dxy2 * cc
            t = torch.where(disc > 0, (b - torch.sqrt(disc.clamp_min(0))) / dxy2, inf)
            z = dz[:, None] * t
            t = torch.where((t > 0.5) & (z >= z0[None, sl]) & (z <= z1[None, sl]), t, inf)
            tmin, arg = t.min(dim=1)
            upd = tmin < best
            best = torch.where(upd, tmin, best)
            lab = torch.where(upd, cl[sl][arg], lab)
    # ---- vertical rectangles ----
    rc = world["rect"]
    mid = 0.5 * (rc[:, 0:2] + rc[:, 2:4])
    rc = rc[np.hypot(mid[:, 0] - x0, mid[:, 1] - y0) < reach + 15]
    if len(rc):
        a_, b_ = to_sensor(rc[:, 0:2]), to_sensor(rc[:, 2:4])
        ax, ay = torch.tensor(a_[:, 0], device=device), torch.tensor(a_[:, 1], device=device)
        ex, ey = torch.tensor(b_[:, 0] - a_[:, 0], device=device), torch.tensor(b_[:, 1] - a_[:, 1], device=device)
        z0 = torch.tensor(rc[:, 4] - SENSOR_H, device=device)
        z1 = torch.tensor(rc[:, 5] - SENSOR_H, device=device)
        cl = torch.tensor(rc[:, 6].astype(np.int64), device=device)
        for k0 in range(0, len(rc), 64):
            sl = slice(k0, k0 + 64)
            den = dx[:, None] * ey[None, sl] - dy[:, None] * ex[None, sl]
            den = torch.where(den.abs() > 1e-9, den, inf)
            t = (ax[None, sl] * ey[None, sl] - ay[None, sl] * ex[None, sl]) / den
            u = (ax[None, sl] * dy[:, None] - ay[None, sl] * dx[:, None]) / den
            z = dz[:, None] * t
            t = torch.where((t > 0.5) & (u >= 0) & (u <= 1) & (z >= z0[None, sl]) & (z <= z1[None, sl]), t, inf)
            tmin, arg = t.min(dim=1)
            upd = tmin < best
            best = torch.where(upd, tmin, best)
            lab = torch.where(upd, cl[sl][arg], lab)
    keep = torch.isfinite(best) & (best < max_range)
    r = best[keep] + (torch.randn(int(keep.sum()), generator=g, dtype=dt) * noise).to(device)
    inten = torch.rand(int(keep.sum()), generator=g, dtype=dt).to(device)
    pts = torch.stack([dx[keep] * r, dy[keep] * r, dz[keep] * r, inten], dim=1).to(torch.float32)
    return pts, lab[keep]


def make_dense_world(seed, n_cyl=2400, radius=62.0):
    """configs[4]-shaped stress scene around the origin: thousands of poles / trunks inside `radius`,
    a few facades; rendered with dense azimuth sampling it yields > 500 instance nodes per scan."""
    rng = np.random.default_rng(seed)
    r = np.sqrt(rng.uniform(4.0 ** 2, radius ** 2, n_cyl))
    a = rng.uniform(-np.pi, np.pi, n_cyl)
    kind = rng.random(n_cyl) < 0.5
    ctr = 500.0   # mid-block of a 1,000 m grid: the whole footprint is terrain
    cyl = np.column_stack([ctr + r * np.cos(a), ctr + r * np.sin(a), np.where(kind, rng.uniform(0.08, 0.16, n_cyl), rng.uniform(0.18, 0.35, n_cyl)),
                           np.zeros(n_cyl), np.where(kind, rng.uniform(3.0, 6.5, n_cyl), rng.uniform(1.8, 3.2, n_cyl)),
                           np.where(kind, POLE, TRUNK)])
    rect = []
    for _ in range(24):
        ang, d, ln = rng.uniform(-np.pi, np.pi), rng.uniform(radius + 3, radius + 25), rng.uniform(10, 30)
        c = np.array([ctr + d * np.cos(ang), ctr + d * np.sin(ang)])
        t = np.array([-np.sin(ang), np.cos(ang)]) * ln / 2
        rect.append((c[0] - t[0], c[1] - t[1], c[0] + t[0], c[1] + t[1], 0.0, rng.uniform(5, 12), BUILDING))
    return dict(poses=np.array([[ctr, ctr, 0.3]]), G=2, block=1000.0, cyl=cyl, rect=np.array(rect, np.float64),
                patches=np.zeros((0, 4)), road_half=3.5, walk=2.0, seed=seed)
