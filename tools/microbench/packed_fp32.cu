// microbenchmark: issue behaviour of packed FP32 (FFMA2/FADD2/FMUL2) on sm_100a
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b) { unsigned r; asm volatile("xor.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ float rcp(float a) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ int f2i(float a) { int r; asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(r) : "f"(a)); return r; }

#define ITER 32768
// MODE 0: 8 scalar FFMA / iter; 1: 8 FFMA2 / iter; 2: 8 FFMA2 + 8 XOR; 3: 8 FFMA + 8 XOR; 4: 8 FADD2; 5: 16 XOR
// 6: 4 FFMA2 + 8 XOR ; 7: 8 MUFU.RCP; 8: 8 F2I; 9: 4 MUFU + 4 F2I; 10: 8 FFMA2 + 4 XOR
template <int MODE>
__global__ void __launch_bounds__(256) kb(float *out, float seed, long long *clk) {
    float f[8]; u64 p[8]; unsigned x[16];
    for (int i = 0; i < 8; ++i) { f[i] = seed + i + threadIdx.x; float2 t = make_float2(f[i], f[i] + 1.f); p[i] = *reinterpret_cast<u64*>(&t); }
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 17 + i;
    float2 cc = make_float2(seed * 0.5f, seed * 0.25f); u64 c2 = *reinterpret_cast<u64*>(&cc);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0 || MODE == 3) f[i] = fma1(f[i], seed, cc.x);
            if (MODE == 1 || MODE == 2 || MODE == 10) p[i] = fma2(p[i], c2, c2);
            if (MODE == 6 && i < 4) p[i] = fma2(p[i], c2, c2);
            if (MODE == 4) p[i] = add2(p[i], c2);
            if (MODE == 2 || MODE == 3 || MODE == 6) x[i] = lop(x[i], x[(i + 1) & 7]);
            if (MODE == 10 && i < 4) x[i] = lop(x[i], x[(i + 1) & 3]);
            if (MODE == 5) { x[i] = lop(x[i], x[(i + 1) & 7]); x[8 + i] = lop(x[8 + i], x[8 + ((i + 1) & 7)]); }
            if (MODE == 7) f[i] = rcp(f[i]);
            if (MODE == 8) x[i] = f2i(f[i]) ^ x[i];
            if (MODE == 9) { if (i < 4) f[i] = rcp(f[i]); else x[i] = f2i(f[i - 4]) ^ x[i]; }
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) { float2 t = *reinterpret_cast<float2*>(&p[i]); s += f[i] + t.x + t.y; }
    unsigned xs = 0; for (int i = 0; i < 16; ++i) xs ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + xs;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int MODE> void run(const char *name, int n_instr, float *out, long long *clk) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int ctas = 148 * 4;   // 4 CTAs x 8 warps = 32 warps / SM = 8 per SMSP
    kb<MODE><<<ctas, 256>>>(out, 1.0001f, clk); cudaDeviceSynchronize();
    cudaEventRecord(a); kb<MODE><<<ctas, 256>>>(out, 1.0001f, clk); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    double cyc_per_iter_smsp = (double)h / ITER;        // 8 warps per SMSP share the issue port
    double clk_ev = ms * 1e-3 * 1.965e9 / ITER;         // clocks per iteration from the event time at 1965 MHz
    printf("%-28s %2d instr/iter/warp: %.2f clk/iter by clock64, %.2f by events (8 warps/SMSP) -> %.3f warp-instr/clk/SMSP   (%.3f ms)\n", name, n_instr,
           cyc_per_iter_smsp, clk_ev, n_instr * 8.0 / clk_ev, ms);
}
int main() {
    float *out; long long *clk; cudaMalloc(&out, 148 * 4 * 256 * 4); cudaMalloc(&clk, 8);
    run<0>("8 FFMA", 8, out, clk);
    run<1>("8 FFMA2", 8, out, clk);
    run<4>("8 FADD2", 8, out, clk);
    run<5>("16 XOR", 16, out, clk);
    run<3>("8 FFMA + 8 XOR", 16, out, clk);
    run<2>("8 FFMA2 + 8 XOR", 16, out, clk);
    run<6>("4 FFMA2 + 8 XOR", 12, out, clk);
    run<10>("8 FFMA2 + 4 XOR", 12, out, clk);
    run<7>("8 MUFU.RCP", 8, out, clk);
    run<8>("8 F2I (+8 XOR)", 16, out, clk);
    run<9>("4 MUFU + 4 F2I (+4 XOR)", 12, out, clk);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
