"""CPU study (no GPU): connected components of the curved-voxel neighbour graph of each DCVC task of a
synthetic street scan -- how much independent work a component-parallel replay would find."""
import sys

import numpy as np

sys.path.insert(0, ".")
from sgtd_b200 import synth, synth_seq  # noqa: E402

w = synth_seq.make_street_world(4541, synth.BASE_SEED + 1)
for si in (0, 1500, 3000):
    p, l = synth_seq.render_at(w, w["poses"][si], 10_000 + si)
    p = p.numpy().astype(np.float64); l = l.numpy()
    for c in (11, 12, 13, 15, 16, 17, 18):
        q = p[l == c]
        if len(q) == 0:
            continue
        rng = np.sqrt((q[:, :3] ** 2).sum(1))
        ok = ~((rng >= 120.0) | (rng <= 0.5))
        pitch = np.degrees(np.arcsin(q[:, 2] / rng)); ang = np.arctan2(q[:, 1], q[:, 0])
        az = np.where(ang > 0, np.degrees(ang), np.degrees(ang + 2 * np.pi))
        mnP = min(0.0, pitch[ok].min()); mxP = max(0.0, pitch[ok].max()); mnR = min(5.0, rng[ok].min()); mxR = max(5.0, rng[ok].max())
        height = int((mxP - mnP) / 1.2)
        b = []; r = mnR; step = 1
        while r <= mxR and len(b) < 1024:
            r += 0.35 - step * 0.0004; b.append(r); step += 1
        b = np.array(b)
        po = np.minimum(np.searchsorted(b, rng, side="right"), len(b) - 1)
        pi = np.round((pitch - mnP) / 1.2).astype(int); ai = np.round(az / 1.2).astype(int)
        po[~ok] = np.minimum(np.searchsorted(b, 0.0, side="right"), len(b) - 1); pi[~ok] = int(round((0 - mnP) / 1.2)); ai[~ok] = 0
        vox = {}
        for k in zip(ai, po, pi):
            vox[k] = vox.get(k, 0) + 1
        keys = list(vox)
        idx = {k: i for i, k in enumerate(keys)}
        parent = list(range(len(keys)))

        def find(x):
            while parent[x] != x:
                parent[x] = parent[parent[x]]; x = parent[x]
            return x
        width = int(round(360.0 / 1.2) + 1)
        for k in keys:
            a0, p0, z0 = k
            for dz in (-1, 0, 1):
                z = z0 + dz
                if z < 0 or z > height: continue
                for dy in (-1, 0, 1):
                    y = p0 + dy
                    if y < 0 or y > len(b): continue
                    for dx in (-1, 0, 1):
                        x = a0 + dx
                        if x < 0: x = width - 1
                        if x > 300: x = 300
                        j = idx.get((x, y, z))
                        if j is not None:
                            ra, rb = find(idx[k]), find(j)
                            if ra != rb: parent[ra] = rb
        comp = {}
        for k in keys:
            r_ = find(idx[k]); comp[r_] = comp.get(r_, 0) + 1
        sizes = sorted(comp.values(), reverse=True)
        print(f"scan {si} cls {c}: npts {len(q)} nvox {len(keys)} comps {len(sizes)} largest {sizes[:5]} share {sizes[0] / len(keys):.2f}")
