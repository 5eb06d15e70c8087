"""Test geometry for scenes larger than the constant-bank limit: a bumpy
tessellated blob and a cloud of small random triangles inside the Cornell box."""
import numpy as np


def blob(center, radius, n_lat, n_lon, bump, seed):
    """Latitude / longitude tessellation of a sphere with a smooth radial bump field:
    2 * n_lon * (n_lat - 1) triangles, counter-clockwise seen from outside."""
    rng = np.random.default_rng(seed)
    k = rng.normal(size=(4, 3)) * 3.0
    ph = rng.uniform(0, 6.28, 4)

    def point(i, j):
        th = np.pi * i / n_lat
        ps = 2 * np.pi * (j % n_lon) / n_lon
        d = np.array([np.sin(th) * np.cos(ps), np.sin(th) * np.sin(ps), np.cos(th)])
        r = radius * (1.0 + bump * sum(np.sin(k[m] @ d + ph[m]) for m in range(4)) / 4.0)
        return np.asarray(center) + r * d
    tris = []
    for i in range(n_lat):
        for j in range(n_lon):
            a, b, c, d = point(i, j), point(i + 1, j), point(i + 1, j + 1), point(i, j + 1)
            if i > 0:
                tris.append([a, b, d])
            if i < n_lat - 1:
                tris.append([b, c, d])
    return np.asarray(tris, np.float32)


def confetti(n, lo, hi, size, seed):
    """n small random triangles with vertices inside [lo, hi]^3-ish bounds."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(lo, hi, size=(n, 1, 3))
    return (c + rng.normal(size=(n, 3, 3)) * size).astype(np.float32)


def big_scene(sv, scene_id, width, height, n_lat=40, n_lon=80, n_confetti=1500):
    """Cornell box `scene_id` + blob (glossy / blue diffuse halves) + confetti (white diffuse)."""
    base = sv.Scene.cornell_box(scene_id, width, height)
    b = blob((0.0, 0.25, 0.15), 0.38, n_lat, n_lon, 0.25, 5)
    c = confetti(n_confetti, (-1.1, -0.9, -1.1), (1.1, 1.1, 1.1), 0.03, 6)
    verts = np.concatenate([b, c])
    mats = np.concatenate([np.where(np.arange(len(b)) % 2 == 0, 2, 8), np.full(len(c), 5)]).astype(np.int32)
    return base.with_triangles(verts, mats), base
