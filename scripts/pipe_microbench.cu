// pipe_microbench.cu — step-0 peaks: issue rates of the instructions the modular arithmetic is made of,
// in lane-ops per clock per SM, measured with register-only loops.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

template <int OP> __device__ __forceinline__ void op(uint32_t& a, uint32_t& b, uint32_t c, double& d, double e) {
    if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(c), "r"(b));            // IMAD
    if (OP == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(c), "r"(b));            // IMAD.HI
    if (OP == 2) { uint64_t t; asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"(a), "r"(c), "l"(((uint64_t)b << 32) | a)); a = (uint32_t)t; b = (uint32_t)(t >> 32); }  // IMAD.WIDE
    if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(c));                             // IADD3
    if (OP == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(c), "r"(b));       // LOP3
    if (OP == 5) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(a) : "r"(b));                  // SHF
    if (OP == 6) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d) : "d"(e));                     // DFMA
    if (OP == 7) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d) : "d"(e));                         // DMUL
    if (OP == 8) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d) : "d"(e));                         // DADD
    if (OP == 9) { float f = __uint_as_float(a); asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f) : "f"(__uint_as_float(c))); a = __float_as_uint(f); }  // FFMA
    if (OP == 10) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(c), "r"(a)); asm volatile("add.u32 %0, %0, %1;" : "+r"(b) : "r"(c)); }       // IMAD + IADD pair
    if (OP == 11) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(c), "r"(a)); asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d) : "d"(e)); } // IMAD + DFMA pair
    if (OP == 12) { asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(c)); asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d) : "d"(e)); }                // IADD + DFMA pair
    if (OP == 13) { int64_t t; asm volatile("cvt.rni.s64.f64 %0, %1;" : "=l"(t) : "d"(d)); d = e + (double)(int32_t)t; }                                     // F2I.S64 + I2F + DADD
    if (OP == 14) asm volatile("cvt.rni.f64.f64 %0, %0;" : "+d"(d));                                                                                          // FRND.F64
    if (OP == 15) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(c), "r"(a)); asm volatile("add.u32 %0, %0, %1;" : "+r"(b) : "r"(c)); asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d) : "d"(e)); }  // IMAD+IADD+DFMA
    if (OP == 16) { uint64_t t = ((uint64_t)b << 32) | a, u = ((uint64_t)c << 32) | c; asm volatile("add.u64 %0, %0, %1;" : "+l"(t) : "l"(u)); a = (uint32_t)t; b = (uint32_t)(t >> 32); }  // 64-bit add
    if (OP == 17) { uint64_t t = ((uint64_t)b << 32) | a, u = ((uint64_t)c << 32) | 12345u, r; asm volatile("mul.hi.u64 %0, %1, %2;" : "=l"(r) : "l"(t), "l"(u)); a = (uint32_t)r; b = (uint32_t)(r >> 32); }  // mul.hi.u64
    if (OP == 19) asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(a) : "r"(b), "r"(c));                                                                      // IDP.4A
    if (OP == 20) asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(a) : "r"(b), "r"(c));                                                                   // IDP.2A
    if (OP == 21) { asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(a) : "r"(b), "r"(c)); asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d) : "d"(e)); }   // IDP.4A + DFMA pair
    if (OP == 18) { uint64_t t = ((uint64_t)b << 32) | a, u = ((uint64_t)c << 32) | 12345u, r; asm volatile("mul.lo.u64 %0, %1, %2;" : "=l"(r) : "l"(t), "l"(u)); a = (uint32_t)r; b = (uint32_t)(r >> 32); }  // mul.lo.u64
}

template <int OP> __global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t c, double e) {
    uint32_t a[CHAINS], b[CHAINS]; double d[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { a[i] = threadIdx.x + i; b[i] = threadIdx.x * 3 + i; d[i] = 1.0 + i + threadIdx.x; }
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) op<OP>(a[i], b[i], c, d[i], e);
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += a[i] + b[i] + (uint32_t)d[i];
    if (s == 0x1234567) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP> void run(const char* name, int per_op, uint32_t* out, int sms, double ghz) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 8;
    k<OP><<<blocks, 256>>>(out, 3u, 1.0000001);
    cudaEventRecord(e0);
    k<OP><<<blocks, 256>>>(out, 3u, 1.0000001);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * 256 * ITERS * 4 * CHAINS * per_op;
    printf("%-28s %8.3f ms  %7.1f lane-ops/clk/SM (at %.3f GHz)\n", name, ms, ops / (ms * 1e-3) / sms / (ghz * 1e9), ghz);
}

// ---- in-warp exchange of 8 doubles per thread: shared memory vs warp shuffles ---------------------------------------------------
// The level-2 NTT moves 8 doubles per thread between two register layouts inside one warp (exchange 2 of ntt.cuh: 256 coefficients
// per warp; lane / register index bits {e7,e1,e0} <-> {e4,e3,e2}).  A: the shared-memory form the kernels use (8 STS.64, __syncwarp,
// 4 LDS.128, padded conflict-free layout).  B: a LOWER BOUND of the shuffle form — the three lane-bit <-> register-bit swaps as three
// rounds of __shfl_xor_sync on half of the registers with the selects around them (the real permutation needs a lane renaming on top).
__global__ void __launch_bounds__(256) exch_smem(double* out, int iters) {
    __shared__ double buf[8 * 288];
    const int t = threadIdx.x, w = t >> 5, L = t & 31;
    double x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = t * 8.0 + k;
    double* base = buf + w * 288;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { const int e = (L >> 2) * 32 + (L & 3) + 4 * k; base[e + 2 * (e >> 4)] = x[k]; }       // pass-2 ownership, pad 2 per 16
        __syncwarp();
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int e = 4 * L + 128 * g; const double2* p = reinterpret_cast<const double2*>(base + e + 2 * (e >> 4));
            const double2 v0 = p[0], v1 = p[1];
            x[4 * g] = v0.x + 1.0; x[4 * g + 1] = v0.y; x[4 * g + 2] = v1.x; x[4 * g + 3] = v1.y;
        }
        __syncwarp();
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k];
    if (s == 12345.0) out[blockIdx.x * blockDim.x + t] = s;
}
__global__ void __launch_bounds__(256) exch_shfl(double* out, int iters) {
    const int t = threadIdx.x, L = t & 31;
    double x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = t * 8.0 + k;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {                      // swap lane bit (1 << r) with register bit (4 >> r)
            const int lb = 1 << r, rb = 4 >> r;
            const bool up = (L & lb) != 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (k & rb) continue;
                const double send = up ? x[k] : x[k | rb];
                const double got = __shfl_xor_sync(0xffffffffu, send, lb);
                if (up) x[k] = got; else x[k | rb] = got;
            }
        }
        x[0] += 1.0;
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k];
    if (s == 12345.0) out[blockIdx.x * blockDim.x + t] = s;
}
void exchange_bench(int sms, double ghz) {
    double* out; cudaMalloc(&out, (size_t)sms * 4 * 256 * 8);
    const int iters = 20000, blocks = sms * 4;
    for (int which = 0; which < 2; ++which) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        if (which == 0) exch_smem<<<blocks, 256>>>(out, iters); else exch_shfl<<<blocks, 256>>>(out, iters);
        cudaEventRecord(e0);
        if (which == 0) exch_smem<<<blocks, 256>>>(out, iters); else exch_shfl<<<blocks, 256>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double per_sm_clk = (ms * 1e-3) * ghz * 1e9 / ((double)iters * 4 * 8);       // cycles per warp-exchange per SM (4 CTAs x 8 warps resident)
        printf("%-44s %8.3f ms  %7.1f SM-cycles per warp exchange of 8 doubles/thread\n",
               which == 0 ? "in-warp exchange via shared memory (as used)" : "in-warp exchange via __shfl_xor (lower bound)", ms, per_sm_clk);
    }
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    printf("%s, %d SMs, clock attr %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
    uint32_t* out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    const int S = p.multiProcessorCount;
    run<0>("IMAD (mad.lo.u32)", 1, out, S, ghz);
    run<1>("IMAD.HI (mad.hi.u32)", 1, out, S, ghz);
    run<2>("IMAD.WIDE (mad.wide.u32)", 1, out, S, ghz);
    run<3>("IADD3 (add.u32)", 1, out, S, ghz);
    run<4>("LOP3", 1, out, S, ghz);
    run<5>("SHF", 1, out, S, ghz);
    run<6>("DFMA", 1, out, S, ghz);
    run<7>("DMUL", 1, out, S, ghz);
    run<8>("DADD", 1, out, S, ghz);
    run<9>("FFMA", 1, out, S, ghz);
    run<10>("IMAD + IADD (2 ops)", 2, out, S, ghz);
    run<11>("IMAD + DFMA (2 ops)", 2, out, S, ghz);
    run<12>("IADD + DFMA (2 ops)", 2, out, S, ghz);
    run<13>("F2I.S64+I2F+DADD (1 unit)", 1, out, S, ghz);
    run<14>("FRND.F64 (cvt.rni)", 1, out, S, ghz);
    run<15>("IMAD+IADD+DFMA (3 ops)", 3, out, S, ghz);
    run<16>("add.u64 (1 unit)", 1, out, S, ghz);
    run<17>("mul.hi.u64 (1 unit)", 1, out, S, ghz);
    run<18>("mul.lo.u64 (1 unit)", 1, out, S, ghz);
    run<19>("IDP.4A (dp4a.s32.s32)", 1, out, S, ghz);
    run<20>("IDP.2A (dp2a.lo)", 1, out, S, ghz);
    run<21>("IDP.4A + DFMA (2 ops)", 2, out, S, ghz);
    exchange_bench(S, ghz);
    return 0;
}
