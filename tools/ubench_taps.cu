// instruction-throughput micro benchmark (one CTA per SM, NW warps): prints warp-instructions / clk / SMSP
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
typedef unsigned long long u64;
#define DEV __device__ __forceinline__
DEV u64 pk(float a, float b){u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;}
DEV u64 fma2(u64 a, u64 b, u64 c){u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;}
DEV u64 mul2(u64 a, u64 b){u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;}
DEV u64 add2(u64 a, u64 b){u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;}
DEV float ffma(float a, float b, float c){float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;}
DEV float fadd(float a, float b){float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r;}
DEV float fmul(float a, float b){float r; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r;}
DEV uint32_t sad(uint32_t a, uint32_t b){uint32_t r; asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(0)); return r;}
DEV uint32_t lea(uint32_t a, uint32_t b){uint32_t r; asm volatile("shl.b32 %0, %1, 7;\n\tadd.u32 %0, %0, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;}
DEV uint32_t prmt(uint32_t a, uint32_t s){uint32_t r; asm volatile("prmt.b32 %0, %1, 0x4B000000, %2;" : "=r"(r) : "r"(a), "r"(s)); return r;}
DEV float lds(uint32_t addr){float r; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr)); return r;}

template <int MODE>
__global__ void __launch_bounds__(1024, 1) bench(int iters, float* out, long long* cyc, int nthreads_active)
{
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 1.0f + 1e-7f * i;
    __syncthreads();
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 4;
    float f[8]; u64 p[8]; uint32_t u[8];
    for (int i = 0; i < 8; ++i) { f[i] = 1.0f + threadIdx.x * 1e-6f + i; p[i] = pk(f[i], f[i] + 1); u[i] = (threadIdx.x * 2654435761u + i * 40503u) & 0x3f3f3f3fu; }   // byte sums <= 252: table index stays inside 64 KB
    const float w = 0.999f + 1e-9f * threadIdx.x; const u64 w2 = pk(w, w + 1e-6f);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) f[i] = ffma(f[i], w, f[(i + 1) & 7]);
                if (MODE == 1) p[i] = fma2(p[i], w2, p[(i + 1) & 7]);
                if (MODE == 2) p[i] = fma2(p[i], pk(w, w), p[(i + 1) & 7]);
                if (MODE == 3) f[i] = fadd(f[i], w);
                if (MODE == 4) p[i] = add2(p[i], w2);
                if (MODE == 5) u[i] = sad(u[i], u[(i + 1) & 7]);
                if (MODE == 6) u[i] = lea(u[i], u[(i + 1) & 7]);
                if (MODE == 7) u[i] = prmt(u[i], 0x7540 + (i & 3));
                if (MODE == 8) f[i] = lds(sbase + ((__float_as_uint(f[i]) >> 3) & 0x3f80));          // lane-private: conflict-free
                if (MODE == 9) {   // old tap: SAD, LEA, LDS, FADD, 3 FFMA   (7 instr)
                    const uint32_t s = u[i] = sad(u[i], u[(i + 1) & 7]);   /* feeds itself: not loop-invariant; <= 1020, then <= 258 */ float ww = lds(lea(s, sbase));
                    f[0] = fadd(f[0], ww); f[1] = ffma(f[5], ww, f[1]); f[2] = ffma(f[6], ww, f[2]); f[3] = ffma(f[7], ww, f[3]);
                }
                if (MODE == 10) {  // new tap: SAD, LEA, LDS, FMUL, 2 FFMA2 (6 instr)
                    const uint32_t s = u[i] = sad(u[i], u[(i + 1) & 7]);   /* feeds itself: not loop-invariant; <= 1020, then <= 258 */ float ww = fmul(lds(lea(s, sbase)), w);
                    p[0] = fma2(p[4], pk(ww, ww), p[0]); p[1] = fma2(p[5], pk(ww, ww), p[1]);
                }
                if (MODE == 11) {  // new tap, folded: SAD, LEA, LDS, 2 FFMA2 (5 instr)
                    const uint32_t s = u[i] = sad(u[i], u[(i + 1) & 7]);   /* feeds itself: not loop-invariant; <= 1020, then <= 258 */ float ww = lds(lea(s, sbase));
                    p[0] = fma2(p[4], pk(ww, ww), p[0]); p[1] = fma2(p[5], pk(ww, ww), p[1]);
                }
                if (MODE == 12) {  // scalar + private table: SAD, LEA, LDS, FMUL, FADD, 3 FFMA (8 instr)
                    const uint32_t s = u[i] = sad(u[i], u[(i + 1) & 7]);   /* feeds itself: not loop-invariant; <= 1020, then <= 258 */ float ww = fmul(lds(lea(s, sbase)), w);
                    f[0] = fadd(f[0], ww); f[1] = ffma(f[5], ww, f[1]); f[2] = ffma(f[6], ww, f[2]); f[3] = ffma(f[7], ww, f[3]);
                }
                if (MODE == 13) {  // 4 independent pixels, scalar: per tap 4x(SAD LEA LDS FADD 3FFMA) -- more ILP
                    const uint32_t s = u[i] = sad(u[i], u[(i + 1) & 7]);   /* feeds itself: not loop-invariant; <= 1020, then <= 258 */ float ww = lds(lea(s, sbase));
                    const int j = i & 1;
                    f[4 * j] = fadd(f[4 * j], ww); f[4 * j + 1] = ffma(w, ww, f[4 * j + 1]); f[4 * j + 2] = ffma(w, ww, f[4 * j + 2]); f[4 * j + 3] = ffma(w, ww, f[4 * j + 3]);
                }
                if (MODE == 14) {  // 2 independent pixels packed
                    const uint32_t s = u[i] = sad(u[i], u[(i + 1) & 7]);   /* feeds itself: not loop-invariant; <= 1020, then <= 258 */ float ww = lds(lea(s, sbase));
                    const int j = i & 1;
                    p[2 * j] = fma2(p[4 + j], pk(ww, ww), p[2 * j]); p[2 * j + 1] = fma2(p[6 + j], pk(ww, ww), p[2 * j + 1]);
                }
                if (MODE == 17) { u[i] = sad(u[i], u[(i + 1) & 7]); f[i] = ffma(f[i], w, f[(i + 1) & 7]); f[(i + 2) & 7] = ffma(f[(i + 2) & 7], w, f[(i + 3) & 7]); }
                if (MODE == 18) { u[i] = sad(u[i], u[(i + 1) & 7]); f[i] = ffma(f[i], w, f[(i + 1) & 7]); f[(i + 2) & 7] = ffma(f[(i + 2) & 7], w, f[(i + 3) & 7]);
                                  f[(i + 4) & 7] = ffma(f[(i + 4) & 7], w, f[(i + 5) & 7]); f[(i + 6) & 7] = ffma(f[(i + 6) & 7], w, f[(i + 7) & 7]); }
                if (MODE == 19) { u[i] = u[i] * 3 + u[(i + 1) & 7]; f[i] = ffma(f[i], w, f[(i + 1) & 7]); f[(i + 2) & 7] = ffma(f[(i + 2) & 7], w, f[(i + 3) & 7]); }
                if (MODE == 20) { u[i] = prmt(u[i], 0x7540 + (i & 3)); f[i] = ffma(f[i], w, f[(i + 1) & 7]); f[(i + 2) & 7] = ffma(f[(i + 2) & 7], w, f[(i + 3) & 7]); }
                if (MODE == 21) { u[i] = lea(u[i], u[(i + 1) & 7]); f[i] = ffma(f[i], w, f[(i + 1) & 7]); }
                if (MODE == 22) { f[i] = lds(sbase + ((__float_as_uint(f[i]) >> 3) & 0x3f80)); f[(i + 2) & 7] = ffma(f[(i + 2) & 7], w, f[(i + 3) & 7]); f[(i + 4) & 7] = ffma(f[(i + 4) & 7], w, f[(i + 5) & 7]);
                                  f[(i + 6) & 7] = ffma(f[(i + 6) & 7], w, f[(i + 7) & 7]); f[(i + 1) & 7] = ffma(f[(i + 1) & 7], w, f[(i + 3) & 7]); }
                if (MODE == 23) { f[i] = lds(sbase + ((__float_as_uint(f[i]) >> 3) & 0x3f80)); u[i] = sad(u[i], u[(i + 1) & 7]); u[(i + 2) & 7] = sad(u[(i + 2) & 7], u[(i + 3) & 7]); }
                if (MODE == 24) f[i] = ffma(f[i], w, f[(i + 2) & 7]);                 // register-bank test: sources two registers apart
                if (MODE == 25) f[i] = ffma(f[i], f[(i + 4) & 7], f[(i + 2) & 7]);   // three sources of one parity
                if (MODE == 26) f[i] = ffma(f[i], f[(i + 3) & 7], f[(i + 2) & 7]);   // two of one parity, one of the other
                if (MODE == 15) f[i] = fmul(f[i], w);
                if (MODE == 16) u[i] = u[i] * 3 + u[(i + 1) & 7];   // IMAD
            }
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p[i])); s += f[i] + a + b + u[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int instr_per_slot, int nw)
{
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    bench<MODE><<<148, nw * 32, 65536>>>(10, out, cyc, 0);
    bench<MODE><<<148, nw * 32, 65536>>>(iters, out, cyc, 0);
    cudaError_t e = cudaDeviceSynchronize();
    long long c[148]; cudaMemcpy(c, cyc, sizeof c, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += c[i]; avg /= 148;
    const double winstr = (double)iters * 32 * instr_per_slot * (nw / 4.0);     // warp-instructions per SMSP
    printf("%-44s warps/SMSP %2d: %6.3f warp-instr/clk/SMSP  (%.2f clk per slot-group per warp) %s\n", name, nw / 4, winstr / avg,
           avg / (iters * 32.0) / (nw / 4.0), e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int nw : {16, 32}) {
        run<0>("FFMA r,r,r", 1, nw); run<1>("FFMA2 (pair weight)", 1, nw); run<2>("FFMA2 (broadcast weight)", 1, nw);
        run<3>("FADD", 1, nw); run<4>("FADD2", 1, nw); run<15>("FMUL", 1, nw); run<16>("IMAD", 1, nw);
        run<5>("VABSDIFF4", 1, nw); run<6>("SHL+ADD (LEA)", 1, nw); run<7>("PRMT", 1, nw); run<8>("LDS lane-private", 1, nw);
        run<9>("tap old: SAD LEA LDS FADD 3FFMA", 7, nw); run<13>("tap old, 2 px interleaved", 7, nw);
        run<11>("tap folded packed: SAD LEA LDS 2FFMA2", 5, nw); run<14>("tap folded packed, 2 px interleaved", 5, nw);
        run<10>("tap private packed: SAD LEA LDS FMUL 2FFMA2", 6, nw);
        run<12>("tap private scalar: SAD LEA LDS FMUL FADD 3FFMA", 8, nw);
        run<24>("FFMA r, same-parity r, w (reuse)", 1, nw); run<25>("FFMA, three sources of one register parity", 1, nw); run<26>("FFMA, sources 2 + 1 by parity", 1, nw);
        run<17>("mix: VABSDIFF4 + 2 FFMA (separate pipes: 3, shared: 4)", 3, nw);
        run<18>("mix: VABSDIFF4 + 4 FFMA (separate: 5, shared: 6)", 5, nw);
        run<19>("mix: IMAD + 2 FFMA", 3, nw);
        run<20>("mix: PRMT + 2 FFMA", 3, nw);
        run<21>("mix: LEA + FFMA", 2, nw);
        run<22>("mix: LDS + SHF/LOP address + 4 FFMA", 5, nw);
        run<23>("mix: LDS + address + 2 VABSDIFF4", 3, nw);
    }
    return 0;
}
