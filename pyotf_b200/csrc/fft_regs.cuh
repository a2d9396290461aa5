// In-register radix-2/4/8/16 butterflies and small complex helpers (sm_100a).
//
// Everything here is templated on the real type T (float / double) and on the transform
// direction DIR: +1 = inverse (numpy ifft kernel e^{+2 pi i k n / N}), -1 = forward.
// Outputs are in natural order; there is no normalisation (callers scale).
//
// These replace the pocketfft passes behind numpy.fft.ifftn / fftn at
// pyotf/otf.py:426, pyotf/phaseretrieval.py:124 and pyotf/utils.py:17,22.
#pragma once
#include <cuda_runtime.h>

namespace pyotf {

// interleaved complex, aligned so that global / shared accesses are single 8 B / 16 B transactions
template <typename T> struct alignas(2 * sizeof(T)) cplx { T x, y; };

template <typename T> struct vec2;
template <> struct vec2<float> { using type = float2; };
template <> struct vec2<double> { using type = double2; };

template <typename T> __device__ __forceinline__ cplx<T> mk(T a, T b) { cplx<T> r; r.x = a; r.y = b; return r; }
template <typename T> __device__ __forceinline__ cplx<T> operator+(cplx<T> a, cplx<T> b) { return mk<T>(a.x + b.x, a.y + b.y); }
template <typename T> __device__ __forceinline__ cplx<T> operator-(cplx<T> a, cplx<T> b) { return mk<T>(a.x - b.x, a.y - b.y); }
#ifndef PYOTF_SCALAR_F32      // -DPYOTF_SCALAR_F32: two FADDs per complex addition (A/B builds)
// Packed FP32 (sm_100a): a complex addition / subtraction is ONE add.f32x2 / sub.f32x2 on the (x, y) register pair --
// the same IEEE result per component as two FADDs at half the issue slots (tools/_dbg/f32x2_bench.cu: both forms run
// at the same 128 FP32 lanes per clock per SM).  The butterflies' additions are a third of the instructions of the
// issue-bound FFT kernels.  Measured (bench.py): K1s 7.01 -> 6.62 us per 128-plane stack (73 -> 77 % of the HBM peak),
// the 1024^2 passes +3 %, everything else unchanged; all parity tests bit-for-bit as before where they compare kernels.
__device__ __forceinline__ cplx<float> operator+(cplx<float> a, cplx<float> b) {
    unsigned long long pa, pb;
    asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b.x), "f"(b.y));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(pa) : "l"(pb));
    cplx<float> r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(pa));
    return r;
}
__device__ __forceinline__ cplx<float> operator-(cplx<float> a, cplx<float> b) {
    unsigned long long pa, pb;
    asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b.x), "f"(b.y));
    asm("sub.rn.f32x2 %0, %0, %1;" : "+l"(pa) : "l"(pb));
    cplx<float> r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(pa));
    return r;
}
#endif
template <typename T> __device__ __forceinline__ cplx<T> cmul(cplx<T> a, cplx<T> b) {
    return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <typename T> __device__ __forceinline__ cplx<T> cmulc(cplx<T> a, cplx<T> b) {  // a * conj(b)
    return mk<T>(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
template <typename T> __device__ __forceinline__ cplx<T> cscale(cplx<T> a, T s) { return mk<T>(a.x * s, a.y * s); }
template <typename T> __device__ __forceinline__ cplx<T> cconj(cplx<T> a) { return mk<T>(a.x, -a.y); }
// multiply by DIR * i  (i.e. e^{DIR * i pi/2})
template <int DIR, typename T> __device__ __forceinline__ cplx<T> mul_i(cplx<T> a) {
    return DIR > 0 ? mk<T>(-a.y, a.x) : mk<T>(a.y, -a.x);
}
// multiply by i^q for runtime q in 0..3 (direction already folded into q)
template <typename T> __device__ __forceinline__ cplx<T> mul_ipow(cplx<T> a, int q) {
    q &= 3;
    if (q == 0) return a;
    if (q == 1) return mk<T>(-a.y, a.x);
    if (q == 2) return mk<T>(-a.x, -a.y);
    return mk<T>(a.y, -a.x);
}

// e^{2 pi i c} for c given in cycles; argument reduced in double so |c| may be large
// (the defocus phase reaches hundreds of cycles, SURVEY.md section 7 hard part 1).
__device__ __forceinline__ void cis_cycles(double c, cplx<float>& out) {
    double f = c - rint(c);            // exact
    float s, co;
    sincospif(2.0f * (float)f, &s, &co);
    out.x = co; out.y = s;
}
__device__ __forceinline__ void cis_cycles(double c, cplx<double>& out) {
    double f = c - rint(c);
    double s, co;
    sincospi(2.0 * f, &s, &co);
    out.x = co; out.y = s;
}

// hypot exactly as glibc >= 2.35 computes it without FMA (Borges, "An improved algorithm for
// hypot(a,b)", the non-__FP_FAST_FMA kernel of sysdeps/ieee754/dbl-64/e_hypot.c) -- which is what
// numpy.hypot returns on x86-64.  The reference's support predicates compare hypot() results with
// thresholds (kr <= na/wl otf.py:355, kr < kmag otf.py:392, r <= 1 zernike.py:322); pixels on
// Pythagorean triples (15, 20 -> 25 bins) land EXACTLY on such a threshold, so kr must be bit-equal
// to numpy's, not merely within an ulp.  Verified bit-for-bit against numpy.hypot on 4e6 random
// pairs and on the fftfreq grids of every BASELINE config (oracle-side check).
// Every operation is an explicitly rounded one: no FMA contraction.
__device__ __forceinline__ double hypot_ref(double x, double y) {
    x = fabs(x); y = fabs(y);
    const double ax = fmax(x, y), ay = fmin(x, y);
    if (ay <= __dmul_rn(ax, 0x1p-53)) return __dadd_rn(ax, ay);
    double h = sqrt(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)));
    double t1, t2;
    if (h <= __dmul_rn(2.0, ay)) {
        const double delta = __dsub_rn(h, ay);
        t1 = __dmul_rn(ax, __dsub_rn(__dmul_rn(2.0, delta), ax));
        t2 = __dmul_rn(__dsub_rn(delta, __dmul_rn(2.0, __dsub_rn(ax, ay))), delta);
    } else {
        const double delta = __dsub_rn(h, ax);
        t1 = __dmul_rn(__dmul_rn(2.0, delta), __dsub_rn(ax, __dmul_rn(2.0, ay)));
        t2 = __dadd_rn(__dmul_rn(__dsub_rn(__dmul_rn(4.0, delta), ay), ay), __dmul_rn(delta, delta));
    }
    h = __dsub_rn(h, __ddiv_rn(__dadd_rn(t1, t2), __dmul_rn(2.0, h)));
    return h;
}

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization
// attribute may start while its predecessor in the stream is still running; it must call pdl_wait() before
// touching anything the predecessor writes, and the predecessor lets it go with pdl_launch_dependents().
// Both are no-ops for ordinary launches.  The PR loop and back-to-back K1 launches use them to overlap a kernel's launch
// latency, prologue and (for K1) its whole compute phase with the tail of the previous kernel.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// bulk asynchronous copies (the TMA engine's non-tensor mode, cp.async.bulk: SASS UBLKCP) and the mbarrier they signal
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
    asm volatile("{\n\t.reg .pred P1;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 :: "r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_and_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <int DIR, typename T> __device__ __forceinline__ void fft2(cplx<T>& a, cplx<T>& b) {
    cplx<T> t = a - b; a = a + b; b = t;
}

// radix-4, natural order in/out
template <int DIR, typename T> __device__ __forceinline__ void fft4(cplx<T>& x0, cplx<T>& x1, cplx<T>& x2, cplx<T>& x3) {
    cplx<T> s02 = x0 + x2, d02 = x0 - x2, s13 = x1 + x3, d13 = mul_i<DIR>(x1 - x3);
    x0 = s02 + s13; x2 = s02 - s13; x1 = d02 + d13; x3 = d02 - d13;
}

// multiply by e^{DIR * 2 pi i * k / 8}, k compile-time
template <int DIR, int K, typename T> __device__ __forceinline__ cplx<T> mul_w8(cplx<T> a) {
    constexpr int k = ((K % 8) + 8) % 8;
    const T h = (T)0.70710678118654752440;
    if (k == 0) return a;
    if (k == 2) return mul_i<DIR>(a);
    if (k == 4) return mk<T>(-a.x, -a.y);
    if (k == 6) return mul_i<-DIR>(a);
    if (k == 1) return DIR > 0 ? mk<T>((a.x - a.y) * h, (a.x + a.y) * h) : mk<T>((a.x + a.y) * h, (a.y - a.x) * h);
    if (k == 3) return DIR > 0 ? mk<T>(-(a.x + a.y) * h, (a.x - a.y) * h) : mk<T>((a.y - a.x) * h, -(a.x + a.y) * h);
    if (k == 5) return DIR > 0 ? mk<T>((a.y - a.x) * h, -(a.x + a.y) * h) : mk<T>(-(a.x + a.y) * h, (a.x - a.y) * h);
    /* k == 7 */ return DIR > 0 ? mk<T>((a.x + a.y) * h, (a.y - a.x) * h) : mk<T>((a.x - a.y) * h, (a.x + a.y) * h);
}

// radix-8 as 2 x radix-4 (decimation in time), natural order in/out
template <int DIR, typename T> __device__ __forceinline__ void fft8(cplx<T>* x) {
    // even / odd split
    cplx<T> e0 = x[0], e1 = x[2], e2 = x[4], e3 = x[6];
    cplx<T> o0 = x[1], o1 = x[3], o2 = x[5], o3 = x[7];
    fft4<DIR>(e0, e1, e2, e3);
    fft4<DIR>(o0, o1, o2, o3);
    o1 = mul_w8<DIR, 1>(o1); o2 = mul_w8<DIR, 2>(o2); o3 = mul_w8<DIR, 3>(o3);
    x[0] = e0 + o0; x[4] = e0 - o0;
    x[1] = e1 + o1; x[5] = e1 - o1;
    x[2] = e2 + o2; x[6] = e2 - o2;
    x[3] = e3 + o3; x[7] = e3 - o3;
}

// multiply by e^{DIR * 2 pi i * k / 16}, k compile-time in 0..15
template <int DIR, int K, typename T> __device__ __forceinline__ cplx<T> mul_w16(cplx<T> a) {
    constexpr int k = ((K % 16) + 16) % 16;
    if (k % 2 == 0) return mul_w8<DIR, k / 2>(a);
    const T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173;
    // odd k: angle = k * 22.5 deg; (cos, sin) from the k=1 and k=3 pair by symmetry
    constexpr int q = k / 4;          // quadrant
    constexpr int r = k % 4;          // 1 or 3
    T c = (r == 1) ? c1 : s1, s = (r == 1) ? s1 : c1;
    // rotate by quadrant: (c, s) -> (c,s), (-s,c), (-c,-s), (s,-c)
    T cc = (q == 0) ? c : (q == 1) ? -s : (q == 2) ? -c : s;
    T ss = (q == 0) ? s : (q == 1) ? c : (q == 2) ? -s : -c;
    if (DIR < 0) ss = -ss;
    return mk<T>(a.x * cc - a.y * ss, a.x * ss + a.y * cc);
}

// radix-16 as 4 x 4, natural order in/out
template <int DIR, typename T> __device__ __forceinline__ void fft16(cplx<T>* x) {
    // k = 4a + b ; out = c + 4d
    cplx<T> t[4][4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        t[b][0] = x[b]; t[b][1] = x[4 + b]; t[b][2] = x[8 + b]; t[b][3] = x[12 + b];
        fft4<DIR>(t[b][0], t[b][1], t[b][2], t[b][3]);   // over a -> c
    }
    // twiddle w16^{b c}
    t[1][1] = mul_w16<DIR, 1>(t[1][1]); t[1][2] = mul_w16<DIR, 2>(t[1][2]); t[1][3] = mul_w16<DIR, 3>(t[1][3]);
    t[2][1] = mul_w16<DIR, 2>(t[2][1]); t[2][2] = mul_w16<DIR, 4>(t[2][2]); t[2][3] = mul_w16<DIR, 6>(t[2][3]);
    t[3][1] = mul_w16<DIR, 3>(t[3][1]); t[3][2] = mul_w16<DIR, 6>(t[3][2]); t[3][3] = mul_w16<DIR, 9>(t[3][3]);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        fft4<DIR>(t[0][c], t[1][c], t[2][c], t[3][c]);   // over b -> d
        x[c] = t[0][c]; x[c + 4] = t[1][c]; x[c + 8] = t[2][c]; x[c + 12] = t[3][c];
    }
}

// radix-16 whose inputs x[4] .. x[11] are known to be zero (band-limited rows: the support box covers less than
// half of the frequencies); the first-stage radix-4 butterflies collapse to two inputs.  x[4..11] are not read.
template <int DIR, typename T> __device__ __forceinline__ void fft16_mid0(cplx<T>* x) {
    cplx<T> t[4][4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const cplx<T> a = x[b], d = x[12 + b], id = mul_i<DIR>(d);
        t[b][0] = a + d; t[b][1] = a - id; t[b][2] = a - d; t[b][3] = a + id;
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

// multiply by e^{DIR * 2 pi i * k / 32}, k compile-time in 0..15
template <int DIR, int K, typename T> __device__ __forceinline__ cplx<T> mul_w32(cplx<T> a) {
    if constexpr (K % 2 == 0) return mul_w16<DIR, K / 2>(a);
    else {
        // cos / sin of k * 11.25 deg for odd k = 1 .. 15
        constexpr double C[8] = {0.98078528040323044913, 0.83146961230254523708, 0.55557023301960222474, 0.19509032201612826785,
                                 -0.19509032201612826785, -0.55557023301960222474, -0.83146961230254523708, -0.98078528040323044913};
        constexpr double S[8] = {0.19509032201612826785, 0.55557023301960222474, 0.83146961230254523708, 0.98078528040323044913,
                                 0.98078528040323044913, 0.83146961230254523708, 0.55557023301960222474, 0.19509032201612826785};
        const T cc = (T)C[K / 2], ss = (T)(DIR > 0 ? S[K / 2] : -S[K / 2]);
        return mk<T>(a.x * cc - a.y * ss, a.x * ss + a.y * cc);
    }
}

// radix-32 as two radix-16 transforms of the even / odd inputs, natural order in/out
template <int DIR, typename T> __device__ __forceinline__ void fft32(cplx<T>* x) {
    cplx<T> e[16], o[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { e[i] = x[2 * i]; o[i] = x[2 * i + 1]; }
    fft16<DIR>(e);
    fft16<DIR>(o);
#define PYOTF_W32(K) { const cplx<T> t = mul_w32<DIR, K>(o[K]); x[K] = e[K] + t; x[K + 16] = e[K] - t; }
    PYOTF_W32(0) PYOTF_W32(1) PYOTF_W32(2) PYOTF_W32(3) PYOTF_W32(4) PYOTF_W32(5) PYOTF_W32(6) PYOTF_W32(7)
    PYOTF_W32(8) PYOTF_W32(9) PYOTF_W32(10) PYOTF_W32(11) PYOTF_W32(12) PYOTF_W32(13) PYOTF_W32(14) PYOTF_W32(15)
#undef PYOTF_W32
}

template <int DIR, int R, typename T> __device__ __forceinline__ void fft_radix(cplx<T>* x) {
    static_assert(R == 1 || R == 2 || R == 4 || R == 8 || R == 16 || R == 32, "radix");
    if constexpr (R == 32) fft32<DIR>(x);
    if constexpr (R == 2) fft2<DIR>(x[0], x[1]);
    if constexpr (R == 4) fft4<DIR>(x[0], x[1], x[2], x[3]);
    if constexpr (R == 8) fft8<DIR>(x);
    if constexpr (R == 16) fft16<DIR>(x);
}

}  // namespace pyotf
