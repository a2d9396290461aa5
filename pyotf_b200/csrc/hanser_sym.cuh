// K1s: HanserPSF._gen_psf + BasePSF.PSFi for the reference's DEFAULT model -- scalar, ideal pupil
// (pyotf/otf.py:362-428 with pupil_base=None, vec_corr="none", and pyotf/otf.py:204-207) -- at N = 256.
// One CTA transforms one whole z-plane on chip; the only HBM traffic is the PSFi store (full 1 KB lines).
//
// pupil * apodisation * defocus depends on (|ky|, |kx|) only and is symmetric under ky <-> kx, so
//   (1) the field is even in y and in x: only the columns kx >= 0 and the rows 0 <= y < 128 are transformed,
//       every computed value lands at (+-y, +-x);
//   (2) output row y = 128 is the transpose of output column x = 128, which the other rows produce anyway.
// A 256-point transform X[c1 + 16 c2] = sum_g w16^{g c2} [ w256^{g c1} sum_q x[16 q + g] w16^{q c1} ] of an
// even sequence needs only 9 + 9 of its 32 radix-16 units: x[16 q + (16 - g)] = x[16 (15 - q) + g] makes the
// first-stage result of residue 16 - g a permutation of residue g's,
//       Z_{16-g}[c1] = conj(w256^{g c1}) * Y_g[-c1],
// and the outputs of second-stage unit 16 - c1 mirror those of unit c1.  The band limit (|f| <= H < 64) leaves
// 7 or 8 of the 16 first-stage inputs non-zero.
//
// Thread mapping: a warp executes ONE unit index (g in stage 1, c1 in stage 2) for 32 columns / 32 rows, one per
// lane, so twiddles and all branch conditions are warp-uniform and every shared-memory access is a unit-stride
// (or odd-stride) conflict-free access.  9 warps (one per unit index) form a group; NG groups per CTA work on
// different 32-row blocks and synchronise with their own named barrier, so the groups drift apart and the
// shared-memory-bound and issue-bound parts of different blocks overlap.
#pragma once
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "hanser.cuh"

namespace pyotf {

// Phase timestamps (debug builds only: -DPYOTF_K1_TIMING, tools/k1_phases.py): 16 marks per (CTA, group)
#ifdef PYOTF_K1_TIMING
static __device__ unsigned long long g_sym_marks[1024 * 2 * 16];
#define SYM_MARK(i) do { if (lane == 0 && (wi == 0 || wi == 9) && blockIdx.y == 0 && blockIdx.x < 1024) g_sym_marks[(blockIdx.x * 2 + grp) * 16 + (i)] = ((i) >= 14 ? k1_globaltimer() : (unsigned long long)clock64()); } while (0)
#else
#define SYM_MARK(i) do { } while (0)
#endif

// radix-16 with inputs 4..11 zero and (X12Z) input 12 zero as well
template <int DIR, bool X12Z, typename T> __device__ __forceinline__ void fft16_band(cplx<T>* x) {
    cplx<T> t[4][4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        if (X12Z && b == 0) {
            t[0][0] = x[0]; t[0][1] = x[0]; t[0][2] = x[0]; t[0][3] = x[0];
        } else {
            const cplx<T> a = x[b], d = x[12 + b], id = mul_i<DIR>(d);
            t[b][0] = a + d; t[b][1] = a - id; t[b][2] = a - d; t[b][3] = a + id;
        }
    }
    t[1][1] = mul_w16<DIR, 1>(t[1][1]); t[1][2] = mul_w16<DIR, 2>(t[1][2]); t[1][3] = mul_w16<DIR, 3>(t[1][3]);
    t[2][1] = mul_w16<DIR, 2>(t[2][1]); t[2][2] = mul_w16<DIR, 4>(t[2][2]); t[2][3] = mul_w16<DIR, 6>(t[2][3]);
    t[3][1] = mul_w16<DIR, 3>(t[3][1]); t[3][2] = mul_w16<DIR, 6>(t[3][2]); t[3][3] = mul_w16<DIR, 9>(t[3][3]);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        fft4<DIR>(t[0][c], t[1][c], t[2][c], t[3][c]);
        x[c] = t[0][c]; x[c + 4] = t[1][c]; x[c + 8] = t[2][c]; x[c + 12] = t[3][c];
    }
}

template <typename T, int NG, int NS = 1> struct SymCfg {
    static constexpr int N = 256;
    static constexpr int NWC = 9 * NG;                        // transform warps: 9 per row group
    static constexpr int NW = NWC + NS * NG, NT = 32 * NW;    // + NS store warps per group
    static constexpr int NHAND = 288 + 32 * NS;               // threads of a FULL / EMPTY hand-over
    static constexpr int NBUF = (sizeof(T) == 4 && NG == 2) ? 2 : 1;   // output staging buffers per group
    static constexpr int EPC = 16 / (int)sizeof(T);           // reals per 16-byte store
    static constexpr int OQ = 128 / EPC + 1;                  // staging: x = EPC * m + j lives at [j][m], m <= 128 / EPC
    static constexpr int OP = EPC * OQ + 1;                   // odd row pitch (reals)
    static constexpr int UP = sizeof(T) == 4 ? 129 : 131;     // odd pitch of U[kx][y], y = 0 .. 128
    static constexpr int ESZ = 9 * 16 * 32;                   // exchange of one group: [c1][g][lane]
    static constexpr int L128N = 132;
    // byte offsets
    __host__ __device__ static size_t u_bytes(int NC) { return sizeof(cplx<T>) * (size_t)NC * UP; }
    __host__ __device__ static size_t e_bytes() { return sizeof(cplx<T>) * (size_t)ESZ * NG; }
    __host__ __device__ static size_t o_bytes() { return sizeof(T) * (size_t)32 * OP * NG * NBUF; }
    __host__ __device__ static size_t base_bytes(int NC) { return u_bytes(NC) + e_bytes() + o_bytes() + sizeof(T) * L128N; }
};

// where the (|fy|, |fx|) table lives while the columns are transformed: inside the (still idle) output staging
// area, else in the tail of U behind column 32 (written only by the second column pass), else appended.
// Returns the byte offset and the total dynamic shared memory.
template <typename T, int NG, int NS> inline void sym_layout(int NC, size_t* t_off, size_t* total) {
    using C = SymCfg<T, NG, NS>;
    const size_t tb = sizeof(cplx<T>) * (size_t)NC * (NC + 1) / 2;      // octant-packed
    const size_t o_off = C::u_bytes(NC) + C::e_bytes();
    if (tb <= C::o_bytes()) { *t_off = o_off; *total = C::base_bytes(NC); return; }
    if (NG == 1 && NC > 32 && tb <= sizeof(cplx<T>) * (size_t)(NC - 32) * C::UP) {
        *t_off = sizeof(cplx<T>) * (size_t)32 * C::UP; *total = C::base_bytes(NC); return;
    }
    *t_off = C::base_bytes(NC); *total = C::base_bytes(NC) + tb;
}

// first stage of an even 256-point inverse transform for residue g (warp-uniform), one transform per lane:
// ld(f) returns the sample of frequency |f| (f = 0 .. H), dst = E + lane receives Z_g[c1] and, for
// 0 < g < 8, Z_{16-g}[c1] (c1 = 0 .. 8) at [(c1 * 16 + g) * 32].
template <bool X12Z, typename T, class LD>
__device__ __forceinline__ void sym_stage1(LD ld, int H, int g, const cplx<T>* tw, cplx<T>* __restrict__ dst) {
    cplx<T> x[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int f = 16 * q + g;                       // frequency 16 q + g
        x[q] = f <= H ? ld(f) : mk<T>(0, 0);
    }
#pragma unroll
    for (int q = X12Z ? 13 : 12; q < 16; ++q) {
        const int f = 256 - 16 * q - g;                 // frequency 16 q + g - 256 = -f
        x[q] = f <= H ? ld(f) : mk<T>(0, 0);
    }
    fft16_band<1, X12Z>(x);
    dst[g * 32] = x[0];
    if (g == 0) {
#pragma unroll
        for (int c = 1; c <= 8; ++c) dst[(c * 16) * 32] = x[c];
        return;
    }
#pragma unroll
    for (int c = 1; c <= 8; ++c) dst[(c * 16 + g) * 32] = cmul(x[c], tw[c - 1]);
    if (g < 8) {
        dst[(16 - g) * 32] = x[0];
#pragma unroll
        for (int c = 1; c <= 8; ++c) dst[(c * 16 + 16 - g) * 32] = cmulc(x[16 - c], tw[c - 1]);
    }
}

// named barriers: 1 + grp = the group's transform warps (288 threads); FULL / EMPTY hand a staging buffer from the
// transform warps to the group's store warp and back (288 + 32 threads)
// (ids are immediates so that ptxas reserves only the barriers in use: two CTAs share an SM)
template <int N> __device__ __forceinline__ void sym_bar_sync(int id) {
    switch (id) {
#define PYOTF_B(I) case I: asm volatile("bar.sync %0, %1;" :: "n"(I), "n"(N) : "memory"); break;
        PYOTF_B(1) PYOTF_B(2) PYOTF_B(3) PYOTF_B(4) PYOTF_B(5) PYOTF_B(6) PYOTF_B(7) PYOTF_B(8) PYOTF_B(9) default: PYOTF_B(10)
#undef PYOTF_B
    }
}
template <int N> __device__ __forceinline__ void sym_bar_arrive(int id) {
    switch (id) {
#define PYOTF_B(I) case I: asm volatile("bar.arrive %0, %1;" :: "n"(I), "n"(N) : "memory"); break;
        PYOTF_B(1) PYOTF_B(2) PYOTF_B(3) PYOTF_B(4) PYOTF_B(5) PYOTF_B(6) PYOTF_B(7) PYOTF_B(8) PYOTF_B(9) default: PYOTF_B(10)
#undef PYOTF_B
    }
}

// DUP: the launch pairs planes at +-z (HanserArgs::npair): a compile-time variant, so that the store loop of the
// ordinary launch -- the bottleneck of the fp32 mode -- carries no extra predicate
template <typename T, int NG, int NS, bool DUP = false>
__global__ void __launch_bounds__(SymCfg<T, NG, NS>::NT, (NG == 1 && sizeof(T) == 4) ? 2 : 1) hanser_sym_kernel(HanserArgs<T> a, int t_off) {
    using C = SymCfg<T, NG, NS>;
    constexpr int NT = C::NT, UP = C::UP, OP = C::OP, OQ = C::OQ, EPC = C::EPC, NBUF = C::NBUF, NPASS = 4 / NG;
    pdl_launch_dependents();
    if (!a.overlap) pdl_wait();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int H = -a.f0, NC = H + 1;
    cplx<T>* const U = reinterpret_cast<cplx<T>*>(smem_raw);
    cplx<T>* const E = reinterpret_cast<cplx<T>*>(smem_raw + C::u_bytes(NC));
    T* const O = reinterpret_cast<T*>(smem_raw + C::u_bytes(NC) + C::e_bytes());
    T* const L128 = O + (size_t)32 * OP * NG * NBUF;
    cplx<T>* const Tt = reinterpret_cast<cplx<T>*>(smem_raw + t_off);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool storer = warp >= C::NWC;                      // warp-uniform role
    const int grp = NG == 1 ? 0 : (storer ? (warp - C::NWC) / NS : warp / 9);
    const int sw = storer ? (warp - C::NWC) % NS : 0;       // which of the group's store warps
    const int wi = storer ? 9 : warp - 9 * grp;             // residue g (first stages) / c1 (second stages)
    // planes at +-z of the default model have the same intensity (the defocus factors are complex conjugates, the pupil
    // is real and even): computed once, stored to both
    const int plane = DUP ? a.pa[blockIdx.x] : blockIdx.x;
    const int plane2 = DUP ? a.pb[blockIdx.x] : plane;
    const bool dup = DUP && plane2 != plane;
    const double z = a.z_in_params ? a.zv[plane] : a.z[plane];
    SYM_MARK(14); SYM_MARK(0);
    const bool x12z = 64 - 8 > H;                           // f = 16*12 + g - 256 is outside the band for every g <= 8
    constexpr int BAR_FULL = 3, BAR_EMPTY = 3 + NG * NBUF;

    // w256^{wi c}, c = 1 .. 8 (fp64-computed table of the handle)
    cplx<T> tw[8];
#pragma unroll
    for (int c = 1; c <= 8; ++c) tw[c - 1] = a.tw[(wi * c) & 255];

    // ---- (|fy|, |fx|) table: apodisation * exp(2 pi i kz z) / N^2, one sincos per octant entry -----------------
    {
        const int ntot = (H / 2 + 1) * (H + 2);
        const T scale = (T)1.0 / ((T)C::N * (T)C::N);
        constexpr int UNR = 3;
        for (int e0 = tid; e0 < ntot; e0 += UNR * NT) {
            double kz[UNR]; T am[UNR]; int ij[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int e = e0 + u * NT;
                const bool ok = e < ntot;
                ij[u] = ok ? a.octij[e] : -1;
                kz[u] = ok ? a.octkz[e] : 0.0;
                am[u] = ok ? a.octamp[e] : (T)0;
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                if (ij[u] >= 0) {
                    const int i = ij[u] >> 16, j = ij[u] & 0xffff;
                    cplx<T> d;
                    cis_cycles(kz[u] * z, d);                   // _calc_defocus   otf.py:357-360
                    d = cscale(d, am[u] * scale);
                    Tt[i * (i + 1) / 2 + j] = d;                // octant-packed: (i, j), i >= j

                }
            }
        }
    }
    __syncthreads();
    SYM_MARK(1);

    // ---- columns: U[kx][y] = sum_ky F(|ky|, kx) w^{ky y}, y = 0 .. 128 -----------------------------------------
    {
        const int nkb = (NC + 31) / 32, npass = (nkb + NG - 1) / NG;
        for (int it = 0; it < npass; ++it) {
            const int kx = 32 * (it * NG + grp) + lane;
            const bool work = !storer && kx < NC;
            cplx<T>* const Eg = E + (size_t)grp * C::ESZ + lane;
            if (work) {
                auto ld = [&](int f) { const int hi = max(f, kx), lo = min(f, kx); return Tt[hi * (hi + 1) / 2 + lo]; };
                if (x12z) sym_stage1<true, T>(ld, H, wi, tw, Eg);
                else sym_stage1<false, T>(ld, H, wi, tw, Eg);
            }
            SYM_MARK(3);
            __syncthreads();
            SYM_MARK(4);
            if (work) {
                cplx<T> x[16];
#pragma unroll
                for (int g = 0; g < 16; ++g) x[g] = Eg[(wi * 16 + g) * 32];
                fft16<1>(x);
                cplx<T>* const Uc = U + (size_t)kx * UP;
                if (wi == 0) {
#pragma unroll
                    for (int c2 = 0; c2 <= 8; ++c2) Uc[16 * c2] = x[c2];
                } else {
#pragma unroll
                    for (int c2 = 0; c2 < 8; ++c2) Uc[wi + 16 * c2] = x[c2];
                    if (wi < 8) {
#pragma unroll
                        for (int c2 = 8; c2 < 16; ++c2) Uc[256 - wi - 16 * c2] = x[c2];
                    }
                }
            }
            SYM_MARK(5);
            __syncthreads();
        }
    }
    SYM_MARK(2);
    T* const psfi = a.psfi + ((size_t)blockIdx.y * a.nz + plane) * (size_t)C::N * C::N;
    T* const psfi2 = a.psfi + ((size_t)blockIdx.y * a.nz + plane2) * (size_t)C::N * C::N;

    if (storer) {
        // ---- store warp of group grp: staged 32-row blocks -> rows y and 256 - y (fftshift-ed positions), 16 bytes
        //      per lane and instruction.  The transform warps never wait for the memory system.
        if (a.overlap == 1) pdl_wait();      // first global store of this launch (it may overwrite a predecessor's output)
        for (int p = 0; p < NPASS; ++p) {
            const int rb = p * NG + grp, buf = p % NBUF;
            const T* const Ob = O + (size_t)(grp * NBUF + buf) * 32 * OP;
            sym_bar_sync<C::NHAND>(BAR_FULL + grp * NBUF + buf);
            if (p == 0) SYM_MARK(10);
            constexpr int NCH = 256 / EPC / 32;          // 16-byte chunks per lane and row
            constexpr int RU = 4;                        // rows in flight: the loop is latency-bound otherwise
            for (int i0 = RU * sw; i0 < 32; i0 += RU * NS) {
                T v[RU][NCH][EPC];
#pragma unroll
                for (int u = 0; u < RU; ++u) {
                    const T* const o = Ob + (i0 + u) * OP;
#pragma unroll
                    for (int k = 0; k < NCH; ++k) {
#pragma unroll
                        for (int j = 0; j < EPC; ++j) {
                            const int xs = (lane + 32 * k) * EPC + j;
                            const int ax = xs < 128 ? 128 - xs : xs - 128;
                            v[u][k][j] = o[(ax % EPC) * OQ + ax / EPC];
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < RU; ++u) {
                    const int yy = 32 * rb + i0 + u;
                    const size_t o0 = (size_t)(yy + 128) * C::N, o1 = (size_t)(128 - yy) * C::N;
#pragma unroll
                    for (int k = 0; k < NCH; ++k) {
                        const int xs = (lane + 32 * k) * EPC;
                        if constexpr (EPC == 4) {
                            const float4 q = make_float4(v[u][k][0], v[u][k][1], v[u][k][2], v[u][k][3]);
                            *reinterpret_cast<float4*>(psfi + o0 + xs) = q;
                            if (yy != 0) *reinterpret_cast<float4*>(psfi + o1 + xs) = q;
                            if (dup) {
                                *reinterpret_cast<float4*>(psfi2 + o0 + xs) = q;
                                if (yy != 0) *reinterpret_cast<float4*>(psfi2 + o1 + xs) = q;
                            }
                        } else {
                            const double2 q = make_double2(v[u][k][0], v[u][k][1]);
                            *reinterpret_cast<double2*>(psfi + o0 + xs) = q;
                            if (yy != 0) *reinterpret_cast<double2*>(psfi + o1 + xs) = q;
                            if (dup) {
                                *reinterpret_cast<double2*>(psfi2 + o0 + xs) = q;
                                if (yy != 0) *reinterpret_cast<double2*>(psfi2 + o1 + xs) = q;
                            }
                        }
                    }
                }
            }
            if (p == 0) SYM_MARK(11);
            if (p + NBUF < NPASS) sym_bar_arrive<C::NHAND>(BAR_EMPTY + grp * NBUF + buf);   // the buffer may be refilled
        }
        SYM_MARK(12);
    } else {
        // corner: out[128][128] = | sum_kx (-1)^kx U[|kx|][128] |^2
        if (warp == 0) {
            cplx<T> s = mk<T>(0, 0);
            for (int kx = lane; kx < NC; kx += 32) {
                const cplx<T> u = U[(size_t)kx * UP + 128];
                const T wgt = (kx == 0 ? (T)1 : (T)2) * ((kx & 1) ? (T)-1 : (T)1);
                s.x += wgt * u.x; s.y += wgt * u.y;
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) { s.x += __shfl_xor_sync(0xffffffffu, s.x, o); s.y += __shfl_xor_sync(0xffffffffu, s.y, o); }
            if (lane == 0) L128[128] = s.x * s.x + s.y * s.y;
        }
        // ---- rows: 32-row blocks, one per group at a time ---------------------------------------------------------
        cplx<T>* const Eg = E + (size_t)grp * C::ESZ + lane;
        for (int p = 0; p < NPASS; ++p) {
            const int rb = p * NG + grp, buf = p % NBUF;
            const int y = 32 * rb + lane;
            if (p > 0) sym_bar_sync<288>(1 + grp);                  // everyone is done reading E
            auto ld = [&](int f) { return U[f * UP + y]; };
            if (x12z) sym_stage1<true, T>(ld, H, wi, tw, Eg);
            else sym_stage1<false, T>(ld, H, wi, tw, Eg);
            if (p == 0) SYM_MARK(6);
            sym_bar_sync<288>(1 + grp);
            if (p == 0) SYM_MARK(7);
            cplx<T> x[16];
#pragma unroll
            for (int g = 0; g < 16; ++g) x[g] = Eg[(wi * 16 + g) * 32];
            fft16<1>(x);
            if (p >= NBUF) sym_bar_sync<C::NHAND>(BAR_EMPTY + grp * NBUF + buf);    // the store warp has drained this buffer
            // |.|^2 at x = wi + 16 c2 (and, mirrored, at 256 - x); staged as [x % EPC][x / EPC]
            T* const o = O + (size_t)(grp * NBUF + buf) * 32 * OP + lane * OP;
            const int lo = (wi % EPC) * OQ + wi / EPC;                      // x = wi + 16 c2, c2 < 8
            const int hi = ((16 - wi) % EPC) * OQ + (16 - wi) / EPC;        // x = 256 - wi - 16 c2 = (16 - wi) + 16 (15 - c2)
#pragma unroll
            for (int c2 = 0; c2 < 8; ++c2) o[lo + (16 / EPC) * c2] = x[c2].x * x[c2].x + x[c2].y * x[c2].y;
            if (wi == 0) {
                const T pw = x[8].x * x[8].x + x[8].y * x[8].y;            // x = 128
                o[128 / EPC] = pw;
                L128[y] = pw;
            } else if (wi < 8) {
#pragma unroll
                for (int c2 = 8; c2 < 16; ++c2) o[hi + (16 / EPC) * (15 - c2)] = x[c2].x * x[c2].x + x[c2].y * x[c2].y;
            }
            if (p == 0) SYM_MARK(8);
            sym_bar_arrive<C::NHAND>(BAR_FULL + grp * NBUF + buf);
        }
        SYM_MARK(13);
    }
    __syncthreads();
    // output row 0 (y = 128): the transpose of column x = 128
    if (warp == C::NWC) {
        for (int xs = lane; xs < 256; xs += 32) {
            const int ax = xs < 128 ? 128 - xs : xs - 128;
            psfi[xs] = L128[ax];
            if (dup) psfi2[xs] = L128[ax];
        }
        // output disjoint from every launch that may still be running: only completion has to stay in stream order
        if (a.overlap == 2) pdl_wait();
    }
    SYM_MARK(15);
}

// Output ranges of the K1s launches per (device, stream) since the last launch that waited for its predecessor before
// storing.  A launch whose output is disjoint from all of them cannot race with a predecessor that is still storing,
// so it need not wait before its own first store: it waits at its very end instead (every CTA, so completion stays
// in stream order and whatever follows in the stream sees all launches done).  A launch that overwrites one of the
// recorded outputs -- or finds the history full -- waits before its first store (which, transitively, means every
// earlier launch has completed) and restarts the history with itself.
constexpr int PYOTF_SYM_HIST = 8;
struct SymStreamHist { int dev = -1; cudaStream_t s = nullptr; const char* lo[PYOTF_SYM_HIST]; const char* hi[PYOTF_SYM_HIST]; int n = 0; };
inline bool sym_output_is_fresh(int dev, cudaStream_t s, const void* out, size_t bytes) {
    static std::mutex mu;
    static SymStreamHist hist[32];
    static int victim = 0;
    std::lock_guard<std::mutex> lk(mu);
    SymStreamHist* h = nullptr;
    for (auto& e : hist) if (e.dev == dev && e.s == s) { h = &e; break; }
    if (!h) { h = &hist[victim]; victim = (victim + 1) % 32; *h = SymStreamHist(); h->dev = dev; h->s = s; }
    const char* lo = (const char*)out; const char* hi = lo + bytes;
    bool fresh = h->n < PYOTF_SYM_HIST;
    for (int i = 0; i < h->n && fresh; ++i) if (lo < h->hi[i] && h->lo[i] < hi) fresh = false;
    if (!fresh) h->n = 0;
    h->lo[h->n] = lo; h->hi[h->n] = hi; ++h->n;
    return fresh;
}

// K1s (hanser_sym.cuh): one CTA per plane, scalar model with the ideal pupil at N = 256, PSFi only
template <typename T, int NG, int NS, bool DUP = false>
int hanser_launch_sym256(const HanserArgs<T>& a, int batch, int smem_limit, cudaStream_t s, bool* done) {
    using C = SymCfg<T, NG, NS>;
    *done = false;
    const int NC = 1 - a.f0;
    size_t t_off = 0, smem = 0;
    sym_layout<T, NG, NS>(NC, &t_off, &smem);
    if (smem > (size_t)smem_limit) return 0;          // the caller falls back to the generic fused kernel
    auto kern = hanser_sym_kernel<T, NG, NS, DUP>;
    static thread_local size_t configured[PYOTF_MAX_DEVICES] = {};
    const int dev_slot = current_device_slot();
    if (smem > configured[dev_slot]) {
        PYOTF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev_slot] = smem;
    }
    HanserArgs<T> aa = a;
    if (aa.overlap) {
        // nothing is read from the predecessor (z by value / caller's promise); if nothing it may still be writing
        // is overwritten either, the launch only has to finish after it
        static const bool no_free = getenv("PYOTF_B200_K1_NO_FREE_OVERLAP") != nullptr;
        int dev = 0;
        cudaGetDevice(&dev);
        const bool fresh = sym_output_is_fresh(dev, s, a.psfi, sizeof(T) * (size_t)batch * a.nz * C::N * C::N);
        if (fresh && !no_free) aa.overlap = 2;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(DUP ? aa.npair : a.nz), (unsigned)batch); cfg.blockDim = dim3(C::NT); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    PYOTF_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, aa, (int)t_off));
    PYOTF_CUDA_OK(cudaGetLastError());
    *done = true;
    return 0;
}

// +-z pairing (host-side zrange only: the positions are known here).  Exact negatives only; z == 0 has no partner.
template <typename T> void sym_pair_planes(HanserArgs<T>& aa) {
    aa.npair = 0;
    // default: on in the fp64 mode (compute-bound: 16.8 -> 13.4 us per cfg1 stack), off in the fp32 mode (already bound by
    // the store stream: 7.6 vs 7.3 us).  PYOTF_B200_K1_ZSYM=1 / 0 forces it (read at every launch: the tests use both).
    const char* e = getenv("PYOTF_B200_K1_ZSYM");
    const bool on = e ? atoi(e) != 0 : sizeof(T) == 8;
    if (!aa.z_in_params || !on || aa.nz > PYOTF_Z_PARAM_MAX) return;
    bool used[PYOTF_Z_PARAM_MAX] = {};
    int np = 0;
    for (int i = 0; i < aa.nz; ++i) {
        if (used[i]) continue;
        used[i] = true;
        int partner = i;
        if (aa.zv[i] != 0.0)
            for (int j = i + 1; j < aa.nz; ++j)
                if (!used[j] && aa.zv[j] == -aa.zv[i]) { partner = j; used[j] = true; break; }
        aa.pa[np] = (unsigned char)i; aa.pb[np] = (unsigned char)partner; ++np;
    }
    if (np < aa.nz) aa.npair = np;                      // at least one pair: launch one CTA per pair / single
}

template <typename T> int hanser_sym_launch(const HanserArgs<T>& a_in, int batch, int smem_limit, cudaStream_t s, bool* done) {
    HanserArgs<T> a = a_in;
    sym_pair_planes(a);
    // one 9-warp group + two store warps per CTA; fp32: two CTAs per SM, so the stores of one plane overlap the
    // transforms of the next.  (Tuning knobs: PYOTF_B200_K1_SYM_NG=2 = two groups in one CTA per SM (fp32),
    // PYOTF_B200_K1_SYM_NS=1 = one store warp.)
    static const int ns = [] { const char* e = getenv("PYOTF_B200_K1_SYM_NS"); return e ? atoi(e) : 2; }();
    if constexpr (sizeof(T) == 4) {
        static const int ng = [] { const char* e = getenv("PYOTF_B200_K1_SYM_NG"); return e ? atoi(e) : 1; }();
        if (ng == 2) return hanser_launch_sym256<T, 2, 1>(a, batch, smem_limit, s, done);
    }
    if (a.npair) return hanser_launch_sym256<T, 1, 2, true>(a, batch, smem_limit, s, done);     // planes at +-z: one CTA per pair
    if (ns == 2) return hanser_launch_sym256<T, 1, 2>(a, batch, smem_limit, s, done);
    return hanser_launch_sym256<T, 1, 1>(a, batch, smem_limit, s, done);
}

}  // namespace pyotf
