#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#define NPRIM 512
__constant__ float4 c_rec[NPRIM];       // (cx,cy,cz,-R2)
__constant__ float4 c_pair[NPRIM / 2];  // (p1a,p1b,p2a,p2b)
__constant__ float2 c_w[NPRIM / 2];     // (wa,wb)
__device__ __forceinline__ float rayc(int r, int k) { return 0.1f * (k + 1) + 0.001f * (threadIdx.x * 4 + r); }

template <int R>
__global__ void __launch_bounds__(128) k_cscalar(unsigned* out, int sweeps) {
    float ux[R], uy[R], uz[R], nou[R];
    for (int r = 0; r < R; r++) { ux[r] = rayc(r, 0); uy[r] = rayc(r, 1); uz[r] = rayc(r, 2); nou[r] = rayc(r, 3); }
    unsigned acc = 0;
    for (int sw = 0; sw < sweeps; sw++) {
        for (int base = 0; base < NPRIM; base += 32) {
            unsigned m[R];
#pragma unroll
            for (int r = 0; r < R; r++) m[r] = 0;
#pragma unroll 8
            for (int j = 0; j < 32; j++) {
                float4 b = c_rec[base + j];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    float a = __fmaf_rn(b.x, ux[r], __fmaf_rn(b.y, uy[r], __fmaf_rn(b.z, uz[r], nou[r])));
                    float d = __fmaf_rn(a, a, b.w);
                    m[r] = __funnelshift_l(__float_as_uint(d), m[r], 1);
                }
            }
#pragma unroll
            for (int r = 0; r < R; r++) acc ^= m[r];
        }
#pragma unroll
        for (int r = 0; r < R; r++) nou[r] += 1e-3f;
    }
    out[blockIdx.x * 128 + threadIdx.x] = acc;
}

template <int R>
__global__ void __launch_bounds__(128) k_cscalar2d(unsigned* out, int sweeps) {
    float u1[R], u2[R], nou[R];
    for (int r = 0; r < R; r++) { u1[r] = rayc(r, 0); u2[r] = rayc(r, 1); nou[r] = rayc(r, 3); }
    unsigned acc = 0;
    for (int sw = 0; sw < sweeps; sw++) {
        for (int base = 0; base < NPRIM; base += 32) {
            unsigned m[R];
#pragma unroll
            for (int r = 0; r < R; r++) m[r] = 0;
#pragma unroll 8
            for (int j = 0; j < 32; j++) {
                float4 b = c_rec[base + j];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    float a = __fmaf_rn(b.x, u1[r], __fmaf_rn(b.y, u2[r], nou[r]));
                    float d = __fmaf_rn(a, a, b.z);
                    m[r] = __funnelshift_l(__float_as_uint(d), m[r], 1);
                }
            }
#pragma unroll
            for (int r = 0; r < R; r++) acc ^= m[r];
        }
#pragma unroll
        for (int r = 0; r < R; r++) nou[r] += 1e-3f;
    }
    out[blockIdx.x * 128 + threadIdx.x] = acc;
}

template <int R>
__global__ void __launch_bounds__(128) k_cpair2d(unsigned* out, int sweeps) {
    float2 u1[R], u2[R], nou[R];
    for (int r = 0; r < R; r++) { u1[r] = make_float2(rayc(r, 0), rayc(r, 0)); u2[r] = make_float2(rayc(r, 1), rayc(r, 1)); nou[r] = make_float2(rayc(r, 3), rayc(r, 3)); }
    unsigned acc = 0;
    for (int sw = 0; sw < sweeps; sw++) {
        for (int base = 0; base < NPRIM / 2; base += 16) {
            unsigned m[R];
#pragma unroll
            for (int r = 0; r < R; r++) m[r] = 0;
#pragma unroll 8
            for (int j = 0; j < 16; j++) {
                float4 A = c_pair[base + j]; float2 B = c_w[base + j];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    float2 a = __ffma2_rn(make_float2(A.x, A.y), u1[r], __ffma2_rn(make_float2(A.z, A.w), u2[r], nou[r]));
                    float2 d = __ffma2_rn(a, a, B);
                    m[r] = __funnelshift_l(__float_as_uint(d.x), m[r], 1);
                    m[r] = __funnelshift_l(__float_as_uint(d.y), m[r], 1);
                }
            }
#pragma unroll
            for (int r = 0; r < R; r++) acc ^= m[r];
        }
#pragma unroll
        for (int r = 0; r < R; r++) nou[r].x += 1e-3f;
    }
    out[blockIdx.x * 128 + threadIdx.x] = acc;
}

template <class F> void timeit(const char* name, double tests_per_thread_per_sweep, int ctas_per_sm, int sms, F launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int sweeps = 400;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0); launch(sms * ctas_per_sm, sweeps); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    double warp_tests = (double) sms * ctas_per_sm * 4 * tests_per_thread_per_sweep * sweeps;
    double cyc = best * 1e-3 * 1.965e9 * sms * 4 / warp_tests;
    printf("%-36s ctas/sm=%d %8.3f ms %7.2f cycles/test %s\n", name, ctas_per_sm, best, cyc, cudaGetLastError() == cudaSuccess ? "" : "ERR");
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    std::vector<float4> h(NPRIM);
    for (int i = 0; i < NPRIM; i++) h[i] = make_float4(0.01f * i, 0.2f, -0.03f * i, -(0.04f + i));
    cudaMemcpyToSymbol(c_rec, h.data(), 16 * NPRIM); cudaMemcpyToSymbol(c_pair, h.data(), 16 * NPRIM / 2); cudaMemcpyToSymbol(c_w, h.data(), 8 * NPRIM / 2);
    unsigned* out; cudaMalloc(&out, 4 * sms * 16 * 128);
    for (int c : {4, 5, 8}) {
        timeit("const scalar R=2", 2.0 * NPRIM, c, sms, [&](int grid, int sw) { k_cscalar<2><<<grid, 128>>>(out, sw); });
        timeit("const scalar R=4", 4.0 * NPRIM, c, sms, [&](int grid, int sw) { k_cscalar<4><<<grid, 128>>>(out, sw); });
        timeit("const 2-D scalar R=2", 2.0 * NPRIM, c, sms, [&](int grid, int sw) { k_cscalar2d<2><<<grid, 128>>>(out, sw); });
        timeit("const 2-D scalar R=4", 4.0 * NPRIM, c, sms, [&](int grid, int sw) { k_cscalar2d<4><<<grid, 128>>>(out, sw); });
        timeit("const 2-D ffma2 pair R=2", 2.0 * NPRIM, c, sms, [&](int grid, int sw) { k_cpair2d<2><<<grid, 128>>>(out, sw); });
        timeit("const 2-D ffma2 pair R=4", 4.0 * NPRIM, c, sms, [&](int grid, int sw) { k_cpair2d<4><<<grid, 128>>>(out, sw); });
    }
    return 0;
}
