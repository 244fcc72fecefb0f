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
    return 0;
}
