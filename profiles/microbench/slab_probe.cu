// Level-1 slab prefilter probe: a = c.u - o.u (3 FMA), disc = a*a - R^2 (1 FMA), sign bit -> mask (1 SHF).
// Variants differ in packing and shared-memory traffic. Reports SMSP cycles per (primitive x ray) test per warp.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#define NPRIM 512

__device__ __forceinline__ float rayc(int r, int k) { return 0.1f * (k + 1) + 0.001f * (threadIdx.x * 4 + r); }

// S3/S4: scalar FFMA, R rays per thread, 16-byte records (cx,cy,cz,-R2)
template <int R>
__global__ void __launch_bounds__(128) k_scalar(const float4* g, unsigned* out, int sweeps) {
    __shared__ float4 s[NPRIM];
    for (int i = threadIdx.x; i < NPRIM; i += 128) s[i] = g[i];
    __syncthreads();
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
                float4 b = s[base + j];
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

// S1: FFMA2 over a ray pair, duplicated 32-byte records (cx,cx,cy,cy)(cz,cz,-R2,-R2); P pairs per thread
template <int P>
__global__ void __launch_bounds__(128) k_raypair(const float4* g2, unsigned* out, int sweeps) {
    __shared__ float4 s[NPRIM * 2];
    for (int i = threadIdx.x; i < NPRIM * 2; i += 128) s[i] = g2[i];
    __syncthreads();
    float2 ux[P], uy[P], uz[P], nou[P];
    for (int p = 0; p < P; p++) {
        ux[p] = make_float2(rayc(2*p, 0), rayc(2*p+1, 0)); uy[p] = make_float2(rayc(2*p, 1), rayc(2*p+1, 1));
        uz[p] = make_float2(rayc(2*p, 2), rayc(2*p+1, 2)); nou[p] = make_float2(rayc(2*p, 3), rayc(2*p+1, 3));
    }
    unsigned acc = 0;
    for (int sw = 0; sw < sweeps; sw++) {
        for (int base = 0; base < NPRIM; base += 32) {
            unsigned m[2 * P];
#pragma unroll
            for (int r = 0; r < 2 * P; r++) m[r] = 0;
#pragma unroll 8
            for (int j = 0; j < 32; j++) {
                float4 A = s[2 * (base + j)], B = s[2 * (base + j) + 1];
#pragma unroll
                for (int p = 0; p < P; p++) {
                    float2 a = __ffma2_rn(make_float2(A.x, A.y), ux[p], __ffma2_rn(make_float2(A.z, A.w), uy[p], __ffma2_rn(make_float2(B.x, B.y), uz[p], nou[p])));
                    float2 d = __ffma2_rn(a, a, make_float2(B.z, B.w));
                    m[2*p] = __funnelshift_l(__float_as_uint(d.x), m[2*p], 1);
                    m[2*p+1] = __funnelshift_l(__float_as_uint(d.y), m[2*p+1], 1);
                }
            }
#pragma unroll
            for (int r = 0; r < 2 * P; r++) acc ^= m[r];
        }
#pragma unroll
        for (int p = 0; p < P; p++) nou[p].x += 1e-3f;
    }
    out[blockIdx.x * 128 + threadIdx.x] = acc;
}

// S2: FFMA2 over a primitive pair, R rays per thread, 32-byte records per PAIR (cx0,cx1,cy0,cy1)(cz0,cz1,-R0,-R1)
template <int R>
__global__ void __launch_bounds__(128) k_primpair(const float4* gp, unsigned* out, int sweeps) {
    __shared__ float4 s[NPRIM];  // NPRIM/2 pairs x 2 float4
    for (int i = threadIdx.x; i < NPRIM; i += 128) s[i] = gp[i];
    __syncthreads();
    float2 ux[R], uy[R], uz[R], nou[R];
    for (int r = 0; r < R; r++) { ux[r] = make_float2(rayc(r, 0), rayc(r, 0)); uy[r] = make_float2(rayc(r, 1), rayc(r, 1)); uz[r] = make_float2(rayc(r, 2), rayc(r, 2)); nou[r] = make_float2(rayc(r, 3), rayc(r, 3)); }
    unsigned acc = 0;
    for (int sw = 0; sw < sweeps; sw++) {
        for (int base = 0; base < NPRIM / 2; base += 16) {   // 16 pairs = 32 prims per mask
            unsigned m[R];
#pragma unroll
            for (int r = 0; r < R; r++) m[r] = 0;
#pragma unroll 8
            for (int j = 0; j < 16; j++) {
                float4 A = s[2 * (base + j)], B = s[2 * (base + j) + 1];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    float2 a = __ffma2_rn(make_float2(A.x, A.y), ux[r], __ffma2_rn(make_float2(A.z, A.w), uy[r], __ffma2_rn(make_float2(B.x, B.y), uz[r], nou[r])));
                    float2 d = __ffma2_rn(a, a, make_float2(B.z, B.w));
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


// S5: 2-D slab (records rotated into the scene basis: a = p1*u1 + p2*u2 - o.u, 2 FMA; disc = a*a - R^2, 1 FMA),
// FFMA2 over a primitive pair, R rays per thread. Per pair: (p1a,p1b,p2a,p2b) in one float4 array, (wa,wb) in a float2 array.
template <int R>
__global__ void __launch_bounds__(128) k_primpair2d(const float4* gp, const float2* gw, unsigned* out, int sweeps) {
    __shared__ float4 s[NPRIM / 2];
    __shared__ float2 w[NPRIM / 2];
    for (int i = threadIdx.x; i < NPRIM / 2; i += 128) { s[i] = gp[i]; w[i] = gw[i]; }
    __syncthreads();
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
                float4 A = s[base + j]; float2 B = w[base + j];
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

// S6: 2-D slab, scalar FFMA, 16-byte records (p1, p2, -R2, pad), R rays per thread
template <int R>
__global__ void __launch_bounds__(128) k_scalar2d(const float4* g, unsigned* out, int sweeps) {
    __shared__ float4 s[NPRIM];
    for (int i = threadIdx.x; i < NPRIM; i += 128) s[i] = g[i];
    __syncthreads();
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
                float4 b = s[base + j];
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

// L: broadcast LDS.128 only
template <int PER>
__global__ void __launch_bounds__(128) k_lds(const float4* g, unsigned* out, int sweeps) {
    __shared__ float4 s[NPRIM * 2];
    for (int i = threadIdx.x; i < NPRIM * 2; i += 128) s[i] = g[i];
    __syncthreads();
    float acc = 0.f;
    for (int sw = 0; sw < sweeps; sw++) {
#pragma unroll 16
        for (int j = 0; j < NPRIM * 2; j++) { float4 v = s[j]; acc += v.x + v.y + v.z + v.w; }
    }
    out[blockIdx.x * 128 + threadIdx.x] = __float_as_uint(acc);
}

template <class F> void timeit(const char* name, double tests_per_thread_per_sweep, int ctas_per_sm, int sms, F launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int sweeps = 400;
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); launch(sms * ctas_per_sm, sweeps); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    double warp_tests = (double) sms * ctas_per_sm * 4 * tests_per_thread_per_sweep * sweeps;  // 4 warps per CTA
    double cyc = best * 1e-3 * 1.965e9 * sms * 4 / warp_tests;
    printf("%-36s ctas/sm=%d %8.3f ms %7.2f cycles/unit %s\n", name, ctas_per_sm, best, cyc, cudaGetLastError() == cudaSuccess ? "" : "ERR");
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    std::vector<float4> h(NPRIM), h2(NPRIM * 2), hp(NPRIM);
    for (int i = 0; i < NPRIM; i++) h[i] = make_float4(0.01f * i, 0.2f, -0.03f * i, -(0.04f + i));
    for (int i = 0; i < NPRIM; i++) { h2[2*i] = make_float4(h[i].x, h[i].x, h[i].y, h[i].y); h2[2*i+1] = make_float4(h[i].z, h[i].z, h[i].w, h[i].w); }
    for (int i = 0; i < NPRIM / 2; i++) { hp[2*i] = make_float4(h[2*i].x, h[2*i+1].x, h[2*i].y, h[2*i+1].y); hp[2*i+1] = make_float4(h[2*i].z, h[2*i+1].z, h[2*i].w, h[2*i+1].w); }
    float4 *g, *g2, *gp; float2* gw; unsigned* out;
    cudaMalloc(&gw, 8 * NPRIM); cudaMemcpy(gw, h.data(), 8 * NPRIM / 2, cudaMemcpyHostToDevice);
    cudaMalloc(&g, 16 * NPRIM); cudaMalloc(&g2, 32 * NPRIM); cudaMalloc(&gp, 16 * NPRIM); cudaMalloc(&out, 4 * sms * 16 * 128);
    cudaMemcpy(g, h.data(), 16 * NPRIM, cudaMemcpyHostToDevice); cudaMemcpy(g2, h2.data(), 32 * NPRIM, cudaMemcpyHostToDevice); cudaMemcpy(gp, hp.data(), 16 * NPRIM, cudaMemcpyHostToDevice);
    printf("%s, %d SMs. slab test: 4 FMA-pipe ops per test -> pipe bound 4.00 cycles/test\n", p.name, sms);
    for (int c : {4, 5, 8}) {
        timeit("scalar R=2 (cycles/test)", 2.0 * NPRIM, c, sms, [&](int grid, int sw) { k_scalar<2><<<grid, 128>>>(g, out, sw); });
        timeit("scalar R=4 (cycles/test)", 4.0 * NPRIM, c, sms, [&](int grid, int sw) { k_scalar<4><<<grid, 128>>>(g, out, sw); });
        timeit("ffma2 ray-pair P=1 (cycles/test)", 2.0 * NPRIM, c, sms, [&](int grid, int sw) { k_raypair<1><<<grid, 128>>>(g2, out, sw); });
        timeit("ffma2 ray-pair P=2 (cycles/test)", 4.0 * NPRIM, c, sms, [&](int grid, int sw) { k_raypair<2><<<grid, 128>>>(g2, out, sw); });
        timeit("ffma2 prim-pair R=2 (cycles/test)", 2.0 * NPRIM, c, sms, [&](int grid, int sw) { k_primpair<2><<<grid, 128>>>(gp, out, sw); });
        timeit("ffma2 prim-pair R=4 (cycles/test)", 4.0 * NPRIM, c, sms, [&](int grid, int sw) { k_primpair<4><<<grid, 128>>>(gp, out, sw); });
        timeit("2-D ffma2 prim-pair R=2 (cycles/test)", 2.0 * NPRIM, c, sms, [&](int grid, int sw) { k_primpair2d<2><<<grid, 128>>>(gp, gw, out, sw); });
        timeit("2-D ffma2 prim-pair R=4 (cycles/test)", 4.0 * NPRIM, c, sms, [&](int grid, int sw) { k_primpair2d<4><<<grid, 128>>>(gp, gw, out, sw); });
        timeit("2-D scalar R=2 (cycles/test)", 2.0 * NPRIM, c, sms, [&](int grid, int sw) { k_scalar2d<2><<<grid, 128>>>(g, out, sw); });
        timeit("2-D scalar R=4 (cycles/test)", 4.0 * NPRIM, c, sms, [&](int grid, int sw) { k_scalar2d<4><<<grid, 128>>>(g, out, sw); });
        timeit("broadcast LDS.128 only (cycles/LDS)", 2.0 * NPRIM, c, sms, [&](int grid, int sw) { k_lds<1><<<grid, 128>>>(g2, out, sw); });
    }
    return 0;
}
