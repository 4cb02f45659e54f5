// Inner-loop probe: variants of the bounding-sphere prefilter sweep, to find what bounds it on B200.
// Each thread carries R rays and sweeps N primitives held in shared memory (or constant memory).
// Output: SMSP cycles per (primitive x ray) test-lane-group, i.e. issue cycles per test per warp.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

#define NPRIM 512
__constant__ float4 c_prims[NPRIM];          // (cx,cy,cz,k)
__constant__ float4 c_prims2[NPRIM * 2];     // duplicated pairs

struct RayF { float dx, dy, dz, nod, m2ox, m2oy, m2oz, oo; };

template <int R> __device__ __forceinline__ void init(RayF (&f)[R]) {
    for (int r = 0; r < R; r++) {
        float t = (float) (threadIdx.x * R + r) * 0.001f;
        f[r].dx = 0.3f + t; f[r].dy = 0.5f - t; f[r].dz = 0.8f; f[r].nod = 0.1f * t; f[r].m2ox = -2.f * t; f[r].m2oy = 1.f; f[r].m2oz = 0.5f; f[r].oo = 3.f + t;
    }
}

// V0: scalar FFMA, prims in smem
template <int R, int UNROLL>
__global__ void __launch_bounds__(128) k_scalar_smem(const float4* g, unsigned* out, int sweeps) {
    __shared__ float4 s[NPRIM];
    for (int i = threadIdx.x; i < NPRIM; i += 128) s[i] = g[i];
    __syncthreads();
    RayF f[R]; init<R>(f);
    unsigned acc = 0;
    for (int sw = 0; sw < sweeps; sw++) {
        for (int base = 0; base < NPRIM; base += 32) {
            unsigned m[R];
#pragma unroll
            for (int r = 0; r < R; r++) m[r] = 0;
#pragma unroll UNROLL
            for (int j = 0; j < 32; j++) {
                float4 b = s[base + j];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    float h = __fmaf_rn(b.x, f[r].dx, __fmaf_rn(b.y, f[r].dy, __fmaf_rn(b.z, f[r].dz, f[r].nod)));
                    float q = __fmaf_rn(b.x, f[r].m2ox, __fmaf_rn(b.y, f[r].m2oy, __fmaf_rn(b.z, f[r].m2oz, b.w + f[r].oo)));
                    float d = __fmaf_rn(h, h, -q);
                    m[r] = __funnelshift_l(__float_as_uint(d), m[r], 1);
                }
            }
#pragma unroll
            for (int r = 0; r < R; r++) acc ^= m[r];
        }
#pragma unroll
        for (int r = 0; r < R; r++) f[r].nod += 1e-3f;
    }
    out[blockIdx.x * 128 + threadIdx.x] = acc;
}

// V1: scalar FFMA, prims in constant memory (uniform operands)
template <int R, int UNROLL>
__global__ void __launch_bounds__(128) k_scalar_const(unsigned* out, int sweeps) {
    RayF f[R]; init<R>(f);
    unsigned acc = 0;
    for (int sw = 0; sw < sweeps; sw++) {
        for (int base = 0; base < NPRIM; base += 32) {
            unsigned m[R];
#pragma unroll
            for (int r = 0; r < R; r++) m[r] = 0;
#pragma unroll UNROLL
            for (int j = 0; j < 32; j++) {
                float4 b = c_prims[base + j];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    float h = __fmaf_rn(b.x, f[r].dx, __fmaf_rn(b.y, f[r].dy, __fmaf_rn(b.z, f[r].dz, f[r].nod)));
                    float q = __fmaf_rn(b.x, f[r].m2ox, __fmaf_rn(b.y, f[r].m2oy, __fmaf_rn(b.z, f[r].m2oz, b.w + f[r].oo)));
                    float d = __fmaf_rn(h, h, -q);
                    m[r] = __funnelshift_l(__float_as_uint(d), m[r], 1);
                }
            }
#pragma unroll
            for (int r = 0; r < R; r++) acc ^= m[r];
        }
#pragma unroll
        for (int r = 0; r < R; r++) f[r].nod += 1e-3f;
    }
    out[blockIdx.x * 128 + threadIdx.x] = acc;
}

// V2: FFMA2 over ray pairs, duplicated prim records in smem (the current product loop). R = 2*P rays.
template <int P, int UNROLL>
__global__ void __launch_bounds__(128) k_pair_smem(const float4* g2, unsigned* out, int sweeps) {
    __shared__ float4 s[NPRIM * 2];
    for (int i = threadIdx.x; i < NPRIM * 2; i += 128) s[i] = g2[i];
    __syncthreads();
    RayF f[2 * P]; init<2 * P>(f);
    float2 dx[P], dy[P], dz[P], nod[P], ax[P], ay[P], az[P], noo[P];
    for (int p = 0; p < P; p++) {
        dx[p] = make_float2(f[2*p].dx, f[2*p+1].dx); dy[p] = make_float2(f[2*p].dy, f[2*p+1].dy); dz[p] = make_float2(f[2*p].dz, f[2*p+1].dz);
        nod[p] = make_float2(f[2*p].nod, f[2*p+1].nod); ax[p] = make_float2(f[2*p].m2ox, f[2*p+1].m2ox); ay[p] = make_float2(f[2*p].m2oy, f[2*p+1].m2oy);
        az[p] = make_float2(f[2*p].m2oz, f[2*p+1].m2oz); noo[p] = make_float2(f[2*p].oo, f[2*p+1].oo);
    }
    unsigned acc = 0;
    for (int sw = 0; sw < sweeps; sw++) {
        for (int base = 0; base < NPRIM; base += 32) {
            unsigned m[2 * P];
#pragma unroll
            for (int r = 0; r < 2 * P; r++) m[r] = 0;
#pragma unroll UNROLL
            for (int j = 0; j < 32; j++) {
                float4 A = s[2 * (base + j)], B = s[2 * (base + j) + 1];
                float2 cx = make_float2(A.x, A.y), cy = make_float2(A.z, A.w), cz = make_float2(B.x, B.y), nk = make_float2(B.z, B.w);
#pragma unroll
                for (int p = 0; p < P; p++) {
                    float2 h = __ffma2_rn(cx, dx[p], __ffma2_rn(cy, dy[p], __ffma2_rn(cz, dz[p], nod[p])));
                    float2 nq = __ffma2_rn(cx, ax[p], __ffma2_rn(cy, ay[p], __ffma2_rn(cz, az[p], __fadd2_rn(nk, noo[p]))));
                    float2 d = __ffma2_rn(h, h, nq);
                    m[2*p] = __funnelshift_l(__float_as_uint(d.x), m[2*p], 1);
                    m[2*p+1] = __funnelshift_l(__float_as_uint(d.y), m[2*p+1], 1);
                }
            }
#pragma unroll
            for (int r = 0; r < 2 * P; r++) acc ^= m[r];
        }
#pragma unroll
        for (int p = 0; p < P; p++) nod[p].x += 1e-3f;
    }
    out[blockIdx.x * 128 + threadIdx.x] = acc;
}

// V3: FFMA2 over ray pairs, duplicated prim records in constant memory
template <int P, int UNROLL>
__global__ void __launch_bounds__(128) k_pair_const(unsigned* out, int sweeps) {
    RayF f[2 * P]; init<2 * P>(f);
    float2 dx[P], dy[P], dz[P], nod[P], ax[P], ay[P], az[P], noo[P];
    for (int p = 0; p < P; p++) {
        dx[p] = make_float2(f[2*p].dx, f[2*p+1].dx); dy[p] = make_float2(f[2*p].dy, f[2*p+1].dy); dz[p] = make_float2(f[2*p].dz, f[2*p+1].dz);
        nod[p] = make_float2(f[2*p].nod, f[2*p+1].nod); ax[p] = make_float2(f[2*p].m2ox, f[2*p+1].m2ox); ay[p] = make_float2(f[2*p].m2oy, f[2*p+1].m2oy);
        az[p] = make_float2(f[2*p].m2oz, f[2*p+1].m2oz); noo[p] = make_float2(f[2*p].oo, f[2*p+1].oo);
    }
    unsigned acc = 0;
    for (int sw = 0; sw < sweeps; sw++) {
        for (int base = 0; base < NPRIM; base += 32) {
            unsigned m[2 * P];
#pragma unroll
            for (int r = 0; r < 2 * P; r++) m[r] = 0;
#pragma unroll UNROLL
            for (int j = 0; j < 32; j++) {
                float4 A = c_prims2[2 * (base + j)], B = c_prims2[2 * (base + j) + 1];
                float2 cx = make_float2(A.x, A.y), cy = make_float2(A.z, A.w), cz = make_float2(B.x, B.y), nk = make_float2(B.z, B.w);
#pragma unroll
                for (int p = 0; p < P; p++) {
                    float2 h = __ffma2_rn(cx, dx[p], __ffma2_rn(cy, dy[p], __ffma2_rn(cz, dz[p], nod[p])));
                    float2 nq = __ffma2_rn(cx, ax[p], __ffma2_rn(cy, ay[p], __ffma2_rn(cz, az[p], __fadd2_rn(nk, noo[p]))));
                    float2 d = __ffma2_rn(h, h, nq);
                    m[2*p] = __funnelshift_l(__float_as_uint(d.x), m[2*p], 1);
                    m[2*p+1] = __funnelshift_l(__float_as_uint(d.y), m[2*p+1], 1);
                }
            }
#pragma unroll
            for (int r = 0; r < 2 * P; r++) acc ^= m[r];
        }
#pragma unroll
        for (int p = 0; p < P; p++) nod[p].x += 1e-3f;
    }
    out[blockIdx.x * 128 + threadIdx.x] = acc;
}

template <class F> void timeit(const char* name, int rays_per_thread, int ctas_per_sm, int sms, F launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int sweeps = 200;
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        launch(sms * ctas_per_sm, sweeps);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    double tests = (double) sms * ctas_per_sm * 128 * rays_per_thread * (double) NPRIM * sweeps;
    double warp_tests = tests / 32.0;                       // (prim x ray) pairs per warp lane-group
    double cyc = best * 1e-3 * 1.965e9 * sms * 4 / warp_tests; // SMSP cycles per test per warp
    printf("%-34s ctas/sm=%d  %8.3f ms  %6.2f cycles/test  %6.1f TFLOP/s(17/test) %s\n", name, ctas_per_sm, best, cyc, 17.0 * tests / (best * 1e-3) / 1e12,
           err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    std::vector<float4> h(NPRIM), h2(NPRIM * 2);
    for (int i = 0; i < NPRIM; i++) {
        h[i] = make_float4(0.01f * i, 0.2f, -0.03f * i, 5.f + i);
        h2[2 * i] = make_float4(h[i].x, h[i].x, h[i].y, h[i].y); h2[2 * i + 1] = make_float4(h[i].z, h[i].z, -h[i].w, -h[i].w);
    }
    float4 *g, *g2; unsigned* out;
    cudaMalloc(&g, sizeof(float4) * NPRIM); cudaMalloc(&g2, sizeof(float4) * NPRIM * 2); cudaMalloc(&out, 4 * sms * 16 * 128);
    cudaMemcpy(g, h.data(), sizeof(float4) * NPRIM, cudaMemcpyHostToDevice); cudaMemcpy(g2, h2.data(), sizeof(float4) * NPRIM * 2, cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(c_prims, h.data(), sizeof(float4) * NPRIM); cudaMemcpyToSymbol(c_prims2, h2.data(), sizeof(float4) * NPRIM * 2);
    printf("%s, %d SMs; ideal FMA-pipe bound = 8.00 cycles/test\n", p.name, sms);
    for (int c : {4, 8}) {
        timeit("scalar smem R=2 unroll8", 2, c, sms, [&](int grid, int sw) { k_scalar_smem<2, 8><<<grid, 128>>>(g, out, sw); });
        timeit("scalar smem R=4 unroll4", 4, c, sms, [&](int grid, int sw) { k_scalar_smem<4, 4><<<grid, 128>>>(g, out, sw); });
        timeit("scalar smem R=4 unroll8", 4, c, sms, [&](int grid, int sw) { k_scalar_smem<4, 8><<<grid, 128>>>(g, out, sw); });
        timeit("scalar const R=2 unroll8", 2, c, sms, [&](int grid, int sw) { k_scalar_const<2, 8><<<grid, 128>>>(out, sw); });
        timeit("scalar const R=4 unroll8", 4, c, sms, [&](int grid, int sw) { k_scalar_const<4, 8><<<grid, 128>>>(out, sw); });
        timeit("ffma2 pair smem P=1 unroll8", 2, c, sms, [&](int grid, int sw) { k_pair_smem<1, 8><<<grid, 128>>>(g2, out, sw); });
        timeit("ffma2 pair smem P=2 unroll4", 4, c, sms, [&](int grid, int sw) { k_pair_smem<2, 4><<<grid, 128>>>(g2, out, sw); });
        timeit("ffma2 pair smem P=2 unroll8", 4, c, sms, [&](int grid, int sw) { k_pair_smem<2, 8><<<grid, 128>>>(g2, out, sw); });
        timeit("ffma2 pair const P=1 unroll8", 2, c, sms, [&](int grid, int sw) { k_pair_const<1, 8><<<grid, 128>>>(out, sw); });
        timeit("ffma2 pair const P=2 unroll8", 4, c, sms, [&](int grid, int sw) { k_pair_const<2, 8><<<grid, 128>>>(out, sw); });
    }
    return 0;
}
