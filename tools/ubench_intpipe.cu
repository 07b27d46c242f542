// Integer-pipe micro-benchmark for sm_100a (B200): measures sustained thread-ops/clk/SM for the
// instructions a BabyBear butterfly is made of. Test infrastructure, not product code.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_intpipe ubench_intpipe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define P 2013265921u
#define ITERS 4096
#define ILP 8

template <int OP>
__global__ void __launch_bounds__(256) kern(uint32_t* out, uint32_t seed, uint32_t w, uint32_t wp) {
    uint32_t a[ILP], b[ILP], wr[4], wpr[4];
#pragma unroll
    for (int i = 0; i < ILP; i++) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed * 3 + i + blockIdx.x; }
#pragma unroll
    for (int i = 0; i < 4; i++) { wr[i] = w + threadIdx.x * 2654435761u + i; wpr[i] = wp ^ (threadIdx.x * 40503u + i); }  // per-thread values
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) { a[i] = a[i] * w + b[i]; }                                   // IMAD
            if (OP == 1) { a[i] = __umulhi(a[i], w) + b[i]; }                          // IMAD.HI.U32
            if (OP == 2) { uint64_t t = (uint64_t)a[i] * w + (((uint64_t)b[i] << 32) | a[i]); a[i] = (uint32_t)t; b[i] = (uint32_t)(t >> 32); } // IMAD.WIDE.U32
            if (OP == 3) { asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])); }   // IADD3
            if (OP == 4) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(w)); } // LOP3
            if (OP == 5) { a[i] = min(a[i] + w, b[i]); }                                // VIADDMNMX
            if (OP == 6) { a[i] = __funnelshift_l(a[i], b[i], 7); }                     // SHF
            if (OP == 7) {  // lazy Shoup DIT butterfly on (a,b): 3 IMAD + 4 ALU
                uint32_t q = __umulhi(b[i], wp);
                uint32_t v = b[i] * w - q * P;
                v = min(v, v - P);
                uint32_t u = min(a[i], a[i] - P);
                a[i] = u + v;
                b[i] = u - v + P;
            }
            if (OP == 8) {  // Montgomery mul a = a*b*R^-1, canonical
                uint64_t t = (uint64_t)a[i] * b[i];
                uint32_t m = (uint32_t)t * 2013265919u;
                uint32_t r = (uint32_t)(t >> 32) - __umulhi(m, P);
                a[i] = min(r, r + P);
            }
            if (OP == 10) { a[i] = a[i] * wr[i & 3] + b[i]; }                          // IMAD, both multiplicands in vector registers
            if (OP == 11) { a[i] = __umulhi(a[i], wr[i & 3]) + b[i]; }                 // IMAD.HI.U32, both multiplicands in vector registers
            if (OP == 12) {  // the same butterfly with a PER-THREAD twiddle (w, w') held in vector registers
                uint32_t q = __umulhi(b[i], wpr[i & 3]);
                uint32_t v = b[i] * wr[i & 3] - q * P;
                v = min(v, v - P);
                uint32_t u = min(a[i], a[i] - P);
                a[i] = u + v;
                b[i] = u - v + P;
            }
            if (OP == 9) {  // butterfly with one add moved to the fma pipe (IMAD x*1+y)
                uint32_t q = __umulhi(b[i], wp);
                uint32_t v = b[i] * w - q * P;
                v = min(v, v - P);
                uint32_t u = min(a[i], a[i] - P);
                uint32_t one; asm volatile("mov.u32 %0, 1;" : "=r"(one));
                asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a[i]) : "r"(u), "r"(one), "r"(v));
                b[i] = u - v + P;
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s ^= a[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static int g_ctas_per_sm = 8;  // 8 CTAs x 8 warps = 16 warps per scheduler; argv[1] overrides (2 = 4 warps per scheduler)
template <int OP>
void run(const char* name, double ops_per_iter, int sms, double mhz) {
    uint32_t* out; int blocks = sms * g_ctas_per_sm, threads = 256;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int r = 0; r < 3; r++) kern<OP><<<blocks, threads>>>(out, 12345u, 1234567u, 2633989657u);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int r = 0; r < reps; r++) kern<OP><<<blocks, threads>>>(out, 12345u, 1234567u, 2633989657u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double total = (double)blocks * threads * ITERS * ILP * ops_per_iter * reps;
    double per_s = total / (ms * 1e-3);
    printf("%-28s %8.3f ms  %8.2f Tthread-op/s  %7.2f thread-ops/clk/SM @%.0f MHz\n", name, ms / reps, per_s / 1e12,
           per_s / sms / (mhz * 1e6), mhz);
    cudaFree(out);
}

int main(int argc, char** argv) {
    if (argc > 1) g_ctas_per_sm = atoi(argv[1]);
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int sms = pr.multiProcessorCount; int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double mhz = khz / 1000.0;
    printf("%s SMs=%d clock=%.0f MHz L2=%d MB smem/SM=%zu, %d warps per scheduler\n", pr.name, sms, mhz, pr.l2CacheSize >> 20, pr.sharedMemPerMultiprocessor, g_ctas_per_sm * 2);
    run<0>("IMAD", 1, sms, mhz);
    run<1>("IMAD.HI.U32", 1, sms, mhz);
    run<2>("IMAD.WIDE.U32", 1, sms, mhz);
    run<3>("IADD3", 1, sms, mhz);
    run<4>("LOP3", 1, sms, mhz);
    run<5>("VIADDMNMX.U32", 1, sms, mhz);
    run<6>("SHF.L", 1, sms, mhz);
    run<7>("butterfly(7 instr) per-instr", 7, sms, mhz);
    run<7>("butterfly(7 instr) per-bfly", 1, sms, mhz);
    run<8>("montmul per-mul", 1, sms, mhz);
    run<9>("butterfly(fma-add) per-bfly", 1, sms, mhz);
    run<10>("IMAD reg*reg", 1, sms, mhz);
    run<11>("IMAD.HI.U32 reg*reg", 1, sms, mhz);
    run<12>("butterfly, per-thread twiddle", 1, sms, mhz);
    return 0;
}
