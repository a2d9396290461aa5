// microbenchmark: scalar FFMA / FADD vs packed FFMA2 / FADD2 issue throughput per SM (sm_100a)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(float* out, int iters, float s) {
    float a[16], b = s, c = 1.0f - s;
    unsigned long long p[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
    unsigned long long bb, cc;
    asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b), "f"(b));
    asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(c), "f"(c));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
            } else if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = a[i] + b;
            } else if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(bb), "l"(cc));
            } else if (MODE == 3) {
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(bb));
            } else if (MODE == 4) {    // mixed: FADD2 + alu-pipe MOV-like (LOP3)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(bb));
                    unsigned lo, hi;
                    asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(p[i]));
                    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "r"(hi), "r"(lo));
                }
            }
        }
    }
    float acc = 0;
    if (MODE < 2) { for (int i = 0; i < 16; ++i) acc += a[i]; }
    else { for (int i = 0; i < 8; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i])); acc += x + y; } }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char* name, int ops_per_inst, int inst_per_iter) {
    float* out; cudaMalloc(&out, 148 * 4 * 1024 * 4);
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps = 4; warps <= 32; warps *= 2) {
        k<MODE><<<148, warps * 32>>>(out, 10, 0.5f);
        cudaEventRecord(e0);
        k<MODE><<<148, warps * 32>>>(out, iters, 0.5f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double winst = (double)iters * 8 * inst_per_iter * warps;           // warp-instructions per SM
        double cyc = ms * 1e-3 * 1.965e9;
        printf("%-28s warps/SM=%2d  %.2f warp-inst/clk/SM  %.1f flop-lanes/clk/SM\n", name, warps, winst / cyc, winst / cyc * 32 * ops_per_inst);
    }
    cudaFree(out);
}
int main() {
    run<0>("FFMA scalar", 1, 16);
    run<1>("FADD scalar", 1, 16);
    run<2>("FFMA2 packed", 2, 8);
    run<3>("FADD2 packed", 2, 8);
    run<4>("FADD2 + swap (2 mov)", 2, 8);
    return 0;
}
