// FFMA vs FFMA2 (fma.rn.f32x2) throughput probe for B200 (sm_100a), with and without
// interleaved ALU-pipe instructions. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>  // 0: FFMA x16, 1: FFMA2 x8 (same flops), 2: FFMA x16 + 4 SHF, 3: FFMA2 x8 + 4 SHF, 4: FFMA2 x8 + 8 SHF
__global__ void __launch_bounds__(256) probe(float* out, float a, float b, int iters) {
    float acc[16];
    unsigned m[8];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = (float) (threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = threadIdx.x * 7 + i;
    float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; it++) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; i++) acc[i] = __fmaf_rn(acc[i], a, b);
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                float2 v = __ffma2_rn(make_float2(acc[2 * i], acc[2 * i + 1]), a2, b2);
                acc[2 * i] = v.x; acc[2 * i + 1] = v.y;
            }
        }
        if (MODE >= 2) {
            const int n = MODE == 4 ? 8 : 4;
#pragma unroll
            for (int i = 0; i < n; i++) m[i] = __funnelshift_l(__float_as_uint(acc[i]), m[i], 1);
        }
    }
    float s = 0.f;
    unsigned ms = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
#pragma unroll
    for (int i = 0; i < 8; i++) ms ^= m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float) ms;
}

template <int MODE> void run(const char* name, float* out, int sms) {
    const int iters = 1 << 14, blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        probe<MODE><<<blocks, threads>>>(out, 0.999f, 0.001f, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    double flops = 2.0 * 16 * (double) iters * blocks * threads;
    printf("%-28s %8.3f ms  %7.2f TFLOP/s\n", name, best, flops / (best * 1e-3) / 1e12);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float* out; cudaMalloc(&out, sizeof(float) * p.multiProcessorCount * 8 * 256);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0>("FFMA x16", out, p.multiProcessorCount);
    run<1>("FFMA2 x8", out, p.multiProcessorCount);
    run<2>("FFMA x16 + 4 SHF", out, p.multiProcessorCount);
    run<3>("FFMA2 x8 + 4 SHF", out, p.multiProcessorCount);
    run<4>("FFMA2 x8 + 8 SHF", out, p.multiProcessorCount);
    return 0;
}
