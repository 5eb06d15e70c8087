"""Prototype check for the merge experiment "true-cell storage with bucket semantics" (profiles/README.md):

The reference's range query visits, for each of the 8 cells n_j around the query, the hash BUCKET
hash(n_j) (src/hashgrid.hxx:124-155) — which also holds far-away aliases, and is visited twice when two
of the 8 coordinates alias. Claim: the multiset of ACCEPTED particles is reproduced by visiting TRUE
cells instead: for each j, scan true cell n_k for every k with hash(n_k) == hash(n_j).

This script builds both structures for random particles and queries with deliberately tiny bucket
counts (so aliasing among the 8 coordinates is frequent) and compares the accepted multisets.
Pure numpy / Python, no GPU, no oracle.  usage: python tools/proto_truecell_merge.py"""
import collections

import numpy as np


def cell_hash(c, n):
    x, y, z = (int(v) & 0xFFFFFFFF for v in c)
    return (((x * 73856093) & 0xFFFFFFFF) ^ ((y * 19349663) & 0xFFFFFFFF) ^ ((z * 83492791) & 0xFFFFFFFF)) % n


def run(seed, num_cells, n_particles=3000, n_queries=1500, radius=0.05):
    rng = np.random.default_rng(seed)
    f32 = np.float32
    p = rng.uniform(0, 1, (n_particles, 3)).astype(f32)
    p[:, 2] = (0.1 * np.sin(6 * p[:, 0])).astype(f32)                       # a surface, like light vertices
    q = (p[rng.integers(0, n_particles, n_queries)] + rng.normal(0, radius * 0.6, (n_queries, 3))).astype(f32)
    bb_min = p.min(0)
    inv = f32(1.0) / (f32(2.0) * f32(radius))
    r2 = f32(radius) * f32(radius)
    coord = np.floor((p - bb_min) * inv).astype(np.int64)
    buckets, cells = collections.defaultdict(list), collections.defaultdict(list)
    for i, c in enumerate(coord):
        buckets[cell_hash(c, num_cells)].append(i)
        cells[tuple(c)].append(i)
    aliased_queries = double_accepts = 0
    for qq in q:
        if np.any(qq - bb_min < 0) or np.any(p.max(0) - qq < 0):
            continue
        cp = (qq - bb_min) * inv
        cf = np.floor(cp)
        base = cf.astype(np.int64)
        other = base + np.where(cp - cf < 0.5, -1, 1)
        n = [tuple(other[a] if (j >> (2 - a)) & 1 else base[a] for a in range(3)) for j in range(8)]
        h = [cell_hash(c, num_cells) for c in n]
        accept = lambda i: np.float32(((qq - p[i]) ** 2).sum(dtype=np.float32)) <= r2
        ref = collections.Counter(i for j in range(8) for i in buckets.get(h[j], ()) if accept(i))
        new = collections.Counter(i for j in range(8) for k in range(8) if h[k] == h[j]
                                  for i in cells.get(n[k], ()) if accept(i))
        assert ref == new, (seed, num_cells, ref - new, new - ref)
        aliased_queries += len(set(h)) < 8
        double_accepts += sum(1 for v in ref.values() if v > 1)
    return aliased_queries, double_accepts


if __name__ == "__main__":
    for num_cells in (37, 64, 512, 4096, 1 << 20):
        a = d = 0
        for seed in range(3):
            x, y = run(seed, num_cells)
            a += x; d += y
        print("buckets %8d: identical accepted multisets; queries with aliased neighbours %5d, particles accepted twice %5d" % (num_cells, a, d))
