"""TEST INFRASTRUCTURE — a numpy walk-through of csrc/knn.cu's algorithm (quantile edges, cell assignment, box growth on
the nearest face, strict termination test), statement for statement, so that the LOGIC of the kernel is checked against
the brute-force oracle on the CPU suite as well.  Not a product path: nothing under fusionsense_b200/ imports it."""
import math

import numpy as np


def build(x, target_per_cell=1.0, max_g=320, max_sample=1 << 18):
    x = np.asarray(x, dtype=np.float32)
    n = len(x)
    g = int(min(max(math.ceil((n / target_per_cell) ** (1.0 / 3.0)), 1), max_g))
    S = min(n, max_sample)
    stride = n // S
    samp = x[np.arange(S) * stride]
    fin = np.isfinite(samp).all(axis=1)
    edges = np.zeros((3, g + 1), dtype=np.float32)
    for a in range(3):
        v = np.sort(samp[fin, a])
        n_fin = len(v)
        for c in range(g + 1):
            r = min((c * n_fin) // g, n_fin - 1)
            edges[a, c] = v[r]
    cells = np.stack([cell_of(edges[a], g, x[:, a]) for a in range(3)], axis=1)
    finite = np.isfinite(x).all(axis=1)
    key = (cells[:, 2].astype(np.int64) * g + cells[:, 1]) * g + cells[:, 0]
    key[~finite] = g ** 3
    order = np.argsort(key, kind="stable")
    cell_start = np.searchsorted(key[order], np.arange(g ** 3 + 1), side="left")
    return {"g": g, "edges": edges, "order": order, "cell_start": cell_start, "x": x}


def cell_of(e, g, p):
    # number of interior edges e[1 .. g-1] that are <= p
    return np.searchsorted(e[1:g], p, side="right").astype(np.int64)


def query_one(ix, q, K, max_steps=96):
    g, edges, order, cs, x = ix["g"], ix["edges"], ix["order"], ix["cell_start"], ix["x"]
    q64 = q.astype(np.float64)
    c = [int(cell_of(edges[a], g, q[a:a + 1])[0]) for a in range(3)]
    best = []  # (d2, idx)

    def scan(b, e):
        for i in order[b:e]:
            d = q64 - x[i].astype(np.float64)
            d2 = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]
            best.append((d2, int(i)))
        best.sort()
        del best[K:]

    lo, hi = list(c), list(c)
    cell = (c[2] * g + c[1]) * g + c[0]
    scan(cs[cell], cs[cell + 1])
    steps = 0
    for step in range(max_steps + 1):
        reach, face = math.inf, -1
        for a in range(3):
            if lo[a] >= 1:
                d = q64[a] - float(edges[a][lo[a]])
                if d < reach:
                    reach, face = d, 2 * a
            if hi[a] + 1 <= g - 1:
                d = float(edges[a][hi[a] + 1]) - q64[a]
                if d < reach:
                    reach, face = d, 2 * a + 1
        if face < 0:
            return best, steps, True
        if len(best) == K and best[K - 1][0] < reach * reach:
            return best, steps, True
        if step == max_steps:
            break
        a = face >> 1
        if face & 1:
            hi[a] += 1
            idx = hi[a]
        else:
            lo[a] -= 1
            idx = lo[a]
        steps += 1
        if a == 0:
            for zz in range(lo[2], hi[2] + 1):
                for yy in range(lo[1], hi[1] + 1):
                    cc = (zz * g + yy) * g + idx
                    scan(cs[cc], cs[cc + 1])
        elif a == 1:
            for zz in range(lo[2], hi[2] + 1):
                row = (zz * g + idx) * g
                scan(cs[row + lo[0]], cs[row + hi[0] + 1])
        else:
            for yy in range(lo[1], hi[1] + 1):
                row = (idx * g + yy) * g
                scan(cs[row + lo[0]], cs[row + hi[0] + 1])
    return best, steps, False
