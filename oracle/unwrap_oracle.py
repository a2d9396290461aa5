"""TEST INFRASTRUCTURE ONLY -- plain-Python restatement of the host phase unwrap in
pyotf_b200/csrc/unwrap.cu (reliability-sorted path following, Herraez et al., Appl. Opt. 41, 7437 (2002);
the algorithm behind skimage.restoration.unwrap_phase, which pyotf calls at phaseretrieval.py:146).

Parity unpinned against skimage (not installed in this image); this file pins the
optimised C++ path (zero-reliability edges joined on the fly, radix sort) to the straightforward
formulation: all edges, one stable sort, one union-find pass.  Small images only (pure Python loops).
"""
import numpy as np

PI = 3.14159265358979323846


def _wrap(x):
    return x - 2 * PI if x > PI else (x + 2 * PI if x < -PI else x)


def _jump(a, b):
    d = a - b
    return 1 if d > PI else (-1 if d < -PI else 0)


def unwrap_phase_2d(image):
    a = np.asarray(image, dtype=np.float64)
    ny, nx = a.shape
    v = [[_wrap(float(a[y, x])) for x in range(nx)] for y in range(ny)]
    rel = [[0.0] * nx for _ in range(ny)]
    for y in range(ny):
        for x in range(nx):
            i = y * nx + x
            if y == 0 or x == 0 or y == ny - 1 or x == nx - 1:
                rel[y][x] = 1e7 + float(((i * 2654435761) & 0xFFFFFFFF) >> 8)   # borders last
                continue
            c = v[y][x]
            H = _wrap(v[y][x - 1] - c) - _wrap(c - v[y][x + 1])
            V = _wrap(v[y - 1][x] - c) - _wrap(c - v[y + 1][x])
            D1 = _wrap(v[y - 1][x - 1] - c) - _wrap(c - v[y + 1][x + 1])
            D2 = _wrap(v[y - 1][x + 1] - c) - _wrap(c - v[y + 1][x - 1])
            rel[y][x] = H * H + V * V + D1 * D1 + D2 * D2
    edges = []
    for y in range(ny):
        for x in range(nx):
            if x + 1 < nx:
                edges.append((rel[y][x] + rel[y][x + 1], y * nx + x, y * nx + x + 1))
            if y + 1 < ny:
                edges.append((rel[y][x] + rel[y + 1][x], y * nx + x, (y + 1) * nx + x))
    edges.sort(key=lambda e: e[0])           # list.sort is stable: ties stay in generation order
    n = ny * nx
    parent, size, pot = list(range(n)), [1] * n, [0] * n
    flat = [v[i // nx][i % nx] for i in range(n)]

    def find(p):
        total = 0
        while parent[p] != p:
            total += pot[p]
            p = parent[p]
        return p, total

    for _, p, q in edges:
        (ra, pa), (rb, pb) = find(p), find(q)
        if ra == rb:
            continue
        d = _jump(flat[p], flat[q]) + pa - pb
        if size[ra] > size[rb]:
            parent[rb], pot[rb] = ra, d
            size[ra] += size[rb]
        else:
            parent[ra], pot[ra] = rb, -d
            size[rb] += size[ra]
    out = np.empty(n)
    for i in range(n):
        out[i] = flat[i] + 2 * PI * find(i)[1]
    return out.reshape(ny, nx)
