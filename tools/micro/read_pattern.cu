// Micro-benchmark: what HBM bandwidth does the tile access pattern of the evaluate kernels reach
// with no other work?  X is feature-major [D][ld]; a CTA reads, per tile of 128 positions, one
// 512-byte segment from each of the D rows (4 MB apart).  Compared with a flat sequential read.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ld_x(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

template <int XB, int DEPTH>
__global__ void __launch_bounds__(128) tile_read(const float *x, size_t ld, int d, int ntiles, float *out) {
    float acc = 0.f;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const float *xp = x + (size_t)tile * 128 + threadIdx.x;
        float buf[DEPTH][XB];
        const int nb = d / XB;
#pragma unroll
        for (int s = 0; s < DEPTH - 1; ++s)
#pragma unroll
            for (int u = 0; u < XB; ++u) buf[s][u] = s < nb ? ld_x(xp + (size_t)(s * XB + u) * ld) : 0.f;
        for (int b = 0; b < nb; b += DEPTH) {
#pragma unroll
            for (int s = 0; s < DEPTH; ++s) {
                const int nxt = b + s + DEPTH - 1;
#pragma unroll
                for (int u = 0; u < XB; ++u)
                    buf[(s + DEPTH - 1) % DEPTH][u] = nxt < nb ? ld_x(xp + (size_t)(nxt * XB + u) * ld) : 0.f;
                if (b + s < nb) {
#pragma unroll
                    for (int u = 0; u < XB; ++u) acc += buf[s][u];
                }
            }
        }
    }
    out[blockIdx.x * 128 + threadIdx.x] = acc;
}

__global__ void flat_read(const float4 *x, size_t n4, float *out) {
    float acc = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(x + i));
        acc += v.x + v.y + v.z + v.w;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
    const int d = 136;
    const size_t n = 1000000, ld = (n + 127) / 128 * 128;
    float *x, *out;
    cudaMalloc(&x, sizeof(float) * d * ld);
    cudaMalloc(&out, sizeof(float) * 148 * 32 * 1024);
    cudaMemset(x, 0, sizeof(float) * d * ld);
    const int ntiles = (int)(n / 128);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    auto time = [&](const char *name, auto launch) {
        launch();
        float best = 1e9f;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(a);
            launch();
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            best = ms < best ? ms : best;
        }
        printf("%-44s %.4f ms  %.0f GB/s\n", name, best, (double)ntiles * 128 * d * 4 / best / 1e6);
    };
    for (int cps : {4, 8, 16}) {
        char nm[128];
        snprintf(nm, sizeof nm, "tile pattern, 8 x 3 blocks in regs, %2d CTA/SM", cps);
        time(nm, [&] { tile_read<8, 3><<<148 * cps, 128>>>(x, ld, d, ntiles, out); });
        snprintf(nm, sizeof nm, "tile pattern, 8 x 5 blocks in regs, %2d CTA/SM", cps);
        time(nm, [&] { tile_read<8, 5><<<148 * cps, 128>>>(x, ld, d, ntiles, out); });
    }
    time("flat sequential float4, 148 x 8 x 256", [&] { flat_read<<<148 * 8, 256>>>((const float4 *)x, (size_t)ntiles * 128 * d / 4, out); });
    time("flat sequential float4, 148 x 16 x 512", [&] { flat_read<<<148 * 16, 512>>>((const float4 *)x, (size_t)ntiles * 128 * d / 4, out); });
    return 0;
}
