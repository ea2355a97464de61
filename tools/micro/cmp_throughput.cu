// Micro-benchmark: how many "compare and count" operations per clock one SM sustains for the
// formulations the ranking kernels could use (lane = candidate, W counters per thread, one
// walked score compared against W held scores).  Build: nvcc -arch=sm_100a -O3 -o cmp cmp_throughput.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define W 16
template <int MODE>
__global__ void k(const double *in, unsigned *out, int iters) {
    double st[W];
    float sf[W];
    unsigned su[W];
    unsigned long long sl[W];
    unsigned cnt[W];
    for (int i = 0; i < W; ++i) {
        st[i] = in[threadIdx.x + 32 * i];
        sf[i] = (float)st[i];
        su[i] = (unsigned)(st[i] * 1e6);
        sl[i] = (unsigned long long)(st[i] * 1e12);
        cnt[i] = 0;
    }
    double sj = in[threadIdx.x + 7];
    const double dj = in[3];
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < W; ++i)
                asm volatile("{ .reg .pred p; setp.ge.f64 p, %1, %2; @p add.u32 %0, %0, 1; }" : "+r"(cnt[i]) : "d"(sj), "d"(st[i]));
            sj += dj;
        } else if (MODE == 1) {
            float fj = (float)sj;
#pragma unroll
            for (int i = 0; i < W; ++i)
                asm volatile("{ .reg .pred p; setp.ge.f32 p, %1, %2; @p add.u32 %0, %0, 1; }" : "+r"(cnt[i]) : "f"(fj), "f"(sf[i]));
            sj += dj;
        } else if (MODE == 2) {
            unsigned uj = (unsigned)it * 2654435761u;
#pragma unroll
            for (int i = 0; i < W; ++i)
                asm volatile("{ .reg .pred p; setp.ge.u32 p, %1, %2; @p add.u32 %0, %0, 1; }" : "+r"(cnt[i]) : "r"(uj), "r"(su[i]));
        } else if (MODE == 3) {
            unsigned long long lj = (unsigned long long)it * 0x9E3779B97F4A7C15ull;
#pragma unroll
            for (int i = 0; i < W; ++i)
                asm volatile("{ .reg .pred p; setp.ge.u64 p, %1, %2; @p add.u32 %0, %0, 1; }" : "+r"(cnt[i]) : "l"(lj), "l"(sl[i]));
        } else if (MODE == 4) {  // DSETP only, predicates folded pairwise (no integer add per compare)
            unsigned acc = 0;
#pragma unroll
            for (int i = 0; i < W; i += 2)
                asm volatile("{ .reg .pred p, q; setp.ge.f64 p, %1, %2; setp.ge.and.f64 q, %1, %3, p; @q add.u32 %0, %0, 1; }"
                             : "+r"(acc) : "d"(sj), "d"(st[i]), "d"(st[i + 1]));
            cnt[0] += acc;
            sj += dj;
        } else if (MODE == 5) {  // f64 compare, add through the FMA pipe (IMAD) instead of the ALU pipe
#pragma unroll
            for (int i = 0; i < W; ++i)
                asm volatile("{ .reg .pred p; setp.ge.f64 p, %1, %2; @p mad.lo.u32 %0, %0, 1, 1; }" : "+r"(cnt[i]) : "d"(sj), "d"(st[i]));
            sj += dj;
        }
    }
    unsigned t = 0;
    for (int i = 0; i < W; ++i) t += cnt[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int MODE>
void run(const char *name, const double *in, unsigned *out, int warps_per_sm) {
    const int iters = 20000, sms = 148;
    const int threads = 128, blocks = sms * warps_per_sm / 4;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k<MODE><<<blocks, threads>>>(in, out, 100);
    cudaEventRecord(a);
    k<MODE><<<blocks, threads>>>(in, out, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double compares = (double)blocks * threads / 32 * iters * W;  // warp-level compare ops
    const double clk = ms * 1e-3 * 1.965e9;
    printf("%-34s warps/SM %2d  %.3f ms  warp-compares per clk per SM %.3f  (per SMSP %.3f)\n", name, warps_per_sm, ms,
           compares / clk / sms, compares / clk / sms / 4);
}

int main() {
    double *in;
    unsigned *out;
    cudaMalloc(&in, 8 * 4096);
    cudaMalloc(&out, 4 * 148 * 64 * 128);
    double h[4096];
    for (int i = 0; i < 4096; ++i) h[i] = (i * 37 % 101) * 0.013;
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    for (int w : {4, 8, 16, 24, 32}) {
        run<0>("f64 DSETP + @p IADD", in, out, w);
        run<5>("f64 DSETP + @p IMAD", in, out, w);
        run<4>("f64 DSETP only (pairs and-ed)", in, out, w);
        run<1>("f32 FSETP + @p IADD", in, out, w);
        run<2>("u32 ISETP + @p IADD", in, out, w);
        run<3>("u64 ISETP x2 + @p IADD", in, out, w);
    }
    return 0;
}
