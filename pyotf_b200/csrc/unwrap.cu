// Host-side 2D phase unwrapping (Herraez, Burton, Lalor, Gdeisat, Appl. Opt. 41, 7437 (2002)):
// reliability-sorted path following with a weighted union-find.  Stand-in for
// skimage.restoration.unwrap_phase at pyotf/phaseretrieval.py:146 (an un-vendored dependency
// of the reference).  It post-processes ONE N x N pupil after the GPU loop has finished
// (SURVEY.md section 8f row N2); the algorithm is a serial sort + union-find, so it runs on
// the host exactly like the reference's.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace {
constexpr double PI = 3.14159265358979323846;
inline double wrap(double x) { return x > PI ? x - 2 * PI : (x < -PI ? x + 2 * PI : x); }
inline int jump(double a, double b) { double d = a - b; return d > PI ? 1 : (d < -PI ? -1 : 0); }

struct UnionFind {
    std::vector<int> parent, size, pot;   // pot[p] = n[p] - n[parent[p]] in units of 2 pi
    explicit UnionFind(int n) : parent(n), size(n, 1), pot(n, 0) { std::iota(parent.begin(), parent.end(), 0); }
    // union by size keeps every tree O(log n) deep, so a plain walk (no path compression, no writes) is fastest
    int find(int p, int& acc) const {     // returns root, acc = n[p] - n[root]
        int total = 0;
        while (parent[p] != p) { total += pot[p]; p = parent[p]; }
        acc = total;
        return p;
    }
};
}  // namespace

extern "C" int pyotf_b200_unwrap_phase_2d(const double* in, double* out, int ny, int nx) {
    PYOTF_REQUIRE(in && out && ny > 0 && nx > 0, "unwrap_phase_2d: bad argument");
    const int n = ny * nx;
    const bool prof = getenv("PYOTF_UNWRAP_PROFILE") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) { if (prof) { auto t1 = std::chrono::steady_clock::now(); fprintf(stderr, "unwrap %s: %.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count()); t0 = t1; } };
    std::vector<double> v(n), rel(n);
    // rows that are entirely zero (the region outside the aperture of a retrieved pupil): where three
    // consecutive rows are, every second difference -- hence the reliability -- is exactly 0
    // per row: the span [first, last] of its non-zero samples (first = nx, last = -1: none).  A pixel whose 3 x 3
    // neighbourhood is entirely zero has every second difference -- hence its reliability -- exactly 0.
    std::vector<int> first(ny), last(ny);
    for (int y = 0; y < ny; ++y) {
        int lo = nx, hi = -1;
        for (int x = 0; x < nx; ++x) {
            const double w = wrap(in[y * nx + x]);        // (np.angle output is already in [-pi, pi])
            v[y * nx + x] = w;
            if (w != 0.0) { if (lo == nx) lo = x; hi = x; }
        }
        first[y] = lo; last[y] = hi;
    }
    for (int y = 0; y < ny; ++y) {
        // columns x with zlo <= x <= zhi may see a non-zero sample of rows y - 1 .. y + 1
        int zlo = nx, zhi = -1;
        if (y > 0 && y < ny - 1) {
            zlo = std::min(first[y - 1], std::min(first[y], first[y + 1])) - 1;
            zhi = std::max(last[y - 1], std::max(last[y], last[y + 1])) + 1;
        }
        for (int x = 0; x < nx; ++x) {
            const int i = y * nx + x;
            if (y == 0 || x == 0 || y == ny - 1 || x == nx - 1) {
                rel[i] = 1e7 + (double)((uint32_t)i * 2654435761u >> 8);    // borders last, deterministic order
                continue;
            }
            if (x < zlo || x > zhi) { rel[i] = 0.0; continue; }
            const double c = v[i];
            const double H = wrap(v[i - 1] - c) - wrap(c - v[i + 1]);
            const double V = wrap(v[i - nx] - c) - wrap(c - v[i + nx]);
            const double D1 = wrap(v[i - nx - 1] - c) - wrap(c - v[i + nx + 1]);
            const double D2 = wrap(v[i - nx + 1] - c) - wrap(c - v[i + nx - 1]);
            rel[i] = H * H + V * V + D1 * D1 + D2 * D2;
        }
    }
    lap("reliability");
    UnionFind uf(n);
    auto join = [&](int a, int b) {
        int pa, pb;
        const int ra = uf.find(a, pa), rb = uf.find(b, pb);
        if (ra == rb) return;
        const int d = jump(v[a], v[b]);                // n[b] - n[a] that makes the pair continuous
        const int nrb_minus_nra = d + pa - pb;
        if (uf.size[ra] > uf.size[rb]) { uf.parent[rb] = ra; uf.pot[rb] = nrb_minus_nra; uf.size[ra] += uf.size[rb]; }
        else { uf.parent[ra] = rb; uf.pot[ra] = -nrb_minus_nra; uf.size[rb] += uf.size[ra]; }
    };
    // Edges are taken in ascending reliability, ties in generation order.  Most edges of a retrieved pupil lie in
    // the zero region outside the aperture (reliability exactly 0): those come first in that order whatever else
    // there is, so they are joined right here as they are generated and only the rest is stored and sorted --
    // the same sequence of joins a stable sort of all edges would give.
    struct Edge { double r; int a, b; };
    std::vector<Edge> edges;
    edges.reserve((size_t)n / 2);
    for (int y = 0; y < ny; ++y)
        for (int x = 0; x < nx; ++x) {
            const int i = y * nx + x;
            if (x + 1 < nx) { const double r = rel[i] + rel[i + 1]; if (r == 0.0) join(i, i + 1); else edges.push_back({r, i, i + 1}); }
            if (y + 1 < ny) { const double r = rel[i] + rel[i + nx]; if (r == 0.0) join(i, i + nx); else edges.push_back({r, i, i + nx}); }
        }
    lap("edges + zero-reliability joins");
    // stable LSD radix sort on the bit patterns of the (positive, finite) reliabilities: the order of
    // std::stable_sort by r at a fraction of its cost; anything else (NaN / inf in the input) takes the comparison sort
    {
        bool radix_ok = true;
        for (const Edge& e : edges) radix_ok = radix_ok && e.r > 0.0 && e.r < 1e300;
        if (edges.empty()) {
        } else if (radix_ok) {
            std::vector<Edge> tmp(edges.size());
            Edge* src = edges.data();
            Edge* dst = tmp.data();
            constexpr int BITS = 11, NB = 1 << BITS;
            auto key = [](const Edge& e) { uint64_t k; memcpy(&k, &e.r, 8); return k; };
            for (int shift = 0; shift < 64; shift += BITS) {
                size_t count[NB] = {};
                for (size_t i = 0; i < edges.size(); ++i) ++count[(key(src[i]) >> shift) & (NB - 1)];
                if (count[(key(src[0]) >> shift) & (NB - 1)] == edges.size()) continue;     // digit constant: no-op pass
                size_t pos = 0;
                for (int d = 0; d < NB; ++d) { const size_t c = count[d]; count[d] = pos; pos += c; }
                for (size_t i = 0; i < edges.size(); ++i) dst[count[(key(src[i]) >> shift) & (NB - 1)]++] = src[i];
                std::swap(src, dst);
            }
            if (src != edges.data()) memcpy(edges.data(), src, edges.size() * sizeof(Edge));
        } else {
            std::stable_sort(edges.begin(), edges.end(), [](const Edge& p, const Edge& q) { return p.r < q.r; });
        }
    }
    lap("sort");
    for (const Edge& e : edges) join(e.a, e.b);
    lap("union");
    for (int i = 0; i < n; ++i) {
        int acc;
        uf.find(i, acc);
        out[i] = v[i] + 2 * PI * (double)acc;
    }
    return 0;
}
