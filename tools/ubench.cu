// tools/ubench.cu -- integer-pipe issue-rate microbenchmark for the DP fill roofline.
// Measures lane-ops/clk/SM of the SASS instructions the fill kernels are made of, with 8
// independent dependency chains per thread so the pipes (not latency) are the limit.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
// Output: one JSON object on stdout (committed as profiles/int_peak_rNN.json).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <string>

#define ITERS 4096
#define ILP 8

template <int OP>
__device__ __forceinline__ unsigned op(unsigned a, unsigned b, unsigned c) {
    if (OP == 0) return a + b;                                  // IADD3 (or IMAD.IADD)
    if (OP == 1) return (unsigned)max((int)a, (int)b);          // VIMNMX s32
    if (OP == 2) return (unsigned)__vimax3_s32((int)a, (int)b, (int)c);   // VIMNMX3
    if (OP == 3) return (unsigned)__viaddmax_s32((int)a, (int)b, (int)c); // VIADDMNMX
    if (OP == 4) return __viaddmax_s16x2(a, b, c);              // VIADDMNMX.S16x2
    if (OP == 5) return __vimax3_s16x2(a, b, c);                // VIMNMX3.S16x2
    if (OP == 6) return __vadd2(a, b);                          // VIADD.16x2
    if (OP == 7) return __byte_perm(a, b, c);                   // PRMT
    if (OP == 8) return (a & b) ^ c;                            // LOP3
    if (OP == 9) return __funnelshift_r(a, b, 2);               // SHF
    if (OP == 10) return a * b + c;                             // IMAD
    if (OP == 11) return ((int)a > (int)b) ? c : a;             // ISETP + SEL
    if (OP == 12) { bool p, q; unsigned r = __vibmax_s16x2(a, b, &p, &q); return r ^ (p ? c : 0u) ^ (q ? b : 0u); } // VIMNMX.S16x2 P,P + 2 SEL-ish
    if (OP == 13) return __vmaxs2(a, b);                        // VIMNMX.S16x2
    if (OP == 14) return (a << 2) + b;                          // LEA
    return a;
}

template <int OP>
__global__ void __launch_bounds__(256) bench(unsigned* out, unsigned seed, unsigned b, unsigned c) {
    unsigned x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = seed + threadIdx.x * 7 + i * 13;
    for (int it = 0; it < ITERS; ++it) {
        unsigned y[ILP];
#pragma unroll
        for (int i = 0; i < ILP; ++i) y[i] = op<OP>(x[i], x[(i + 1) % ILP], x[(i + 3) % ILP]);
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = y[i];
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s ^= x[i];
    if (s == 0x12345678u) out[0] = s;
}

// shuffle throughput
__global__ void __launch_bounds__(256) bench_shfl(unsigned* out, unsigned seed) {
    unsigned x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = seed + threadIdx.x * 7 + i * 13;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = __shfl_up_sync(0xffffffffu, x[i], 1);
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s ^= x[i];
    if (s == 0x12345678u) out[0] = s;
}

// the mix of the s16x2 fill inner loop: 2 VIADDMNMX + LOP3 + VIADD + PRMT + LOP3 + SHF/LEA
__global__ void __launch_bounds__(256) bench_mix(unsigned* out, unsigned seed, unsigned b, unsigned c) {
    unsigned x[ILP], acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = seed + threadIdx.x * 7 + i * 13;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            unsigned sel = x[i] ^ b;
            unsigned s = __byte_perm(b, c, sel);
            unsigned z = __viaddmax_s16x2(x[i], s, x[(i + 1) % ILP]);
            z = __viaddmax_s16x2(x[(i + 3) % ILP], c, z);
            acc = __funnelshift_r(acc, z, 2);
            x[i] = __vadd2(z & 0xfffcfffcu, b);
        }
    }
    unsigned s = acc;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s ^= x[i];
    if (s == 0x12345678u) out[0] = s;
}

template <class F>
static double run(F launch, int sms, double ops_per_thread_iter, const cudaDeviceProp& prop, double* mhz_out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double threads = (double)sms * 8 * 256;
    const double ops = threads * ITERS * ops_per_thread_iter;
    *mhz_out = 0;
    return ops / (best * 1e-3);  // lane-ops per second
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    unsigned* d; cudaMalloc(&d, 64);
    const char* names[] = {"IADD3", "VIMNMX.s32", "VIMNMX3.s32", "VIADDMNMX.s32", "VIADDMNMX.S16x2", "VIMNMX3.S16x2",
                           "VIADD.16x2", "PRMT", "LOP3", "SHF", "IMAD", "ISETP+SEL", "VIMNMX.S16x2.P+2SEL", "VIMNMX.S16x2", "LEA"};
    double clk_hz = 0; int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0); clk_hz = khz * 1e3;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_rate_attr_mhz\": %.0f, \"results\": {", prop.name, sms, clk_hz / 1e6);
    dim3 grid(sms * 8), block(256);
    double mhz;
#define RUN(OP) { double r = run([&] { bench<OP><<<grid, block>>>(d, 1u, 3u, 5u); }, sms, ILP, prop, &mhz); \
    printf("%s\"%s\": {\"Tops\": %.2f, \"per_clk_per_sm_at_attr_clock\": %.1f}", OP ? ", " : "", names[OP], r / 1e12, r / clk_hz / sms); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) RUN(14)
    { double r = run([&] { bench_shfl<<<grid, block>>>(d, 1u); }, sms, ILP, prop, &mhz);
      printf(", \"SHFL.UP\": {\"Tops\": %.2f, \"per_clk_per_sm_at_attr_clock\": %.1f}", r / 1e12, r / clk_hz / sms); }
    { double r = run([&] { bench_mix<<<grid, block>>>(d, 1u, 3u, 5u); }, sms, ILP * 7, prop, &mhz);
      printf(", \"fill_mix_7ops\": {\"Tops\": %.2f, \"per_clk_per_sm_at_attr_clock\": %.1f}", r / 1e12, r / clk_hz / sms); }
    printf("}}\n");
    return 0;
}
