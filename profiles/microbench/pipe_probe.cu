// What bounds the 2-D slab inner loop (3 FFMA2 + 2 SHF per ray per primitive pair, operands from constant bank)?
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#define NPRIM 512
__constant__ float4 c_pair[NPRIM / 2];  // (p1a,p1b,p2a,p2b)
__constant__ float2 c_w[NPRIM / 2];     // (wa,wb)
__device__ __forceinline__ float rayc(int r, int k) { return 0.1f * (k + 1) + 0.001f * (threadIdx.x * 4 + r); }

// FLAGS: bit0 = no SHF (sum instead, 1 FADD2 per pair... replaced by xor of raw bits every 8), bit1 = no LDCU (same record), bit2 = SHF via LOP3-free accumulate (use min)
template <int R, int FLAGS>
__global__ void __launch_bounds__(128) k_cpair2d(unsigned* out, int sweeps) {
    float2 u1[R], u2[R], nou[R];
    for (int r = 0; r < R; r++) { u1[r] = make_float2(rayc(r, 0), rayc(r, 0)); u2[r] = make_float2(rayc(r, 1), rayc(r, 1)); nou[r] = make_float2(rayc(r, 3), rayc(r, 3)); }
    unsigned acc = 0;
    for (int sw = 0; sw < sweeps; sw++) {
        for (int base = 0; base < NPRIM / 2; base += 16) {
            unsigned m[R];
            float2 fm[R];
#pragma unroll
            for (int r = 0; r < R; r++) { m[r] = 0; fm[r] = make_float2(0.f, 0.f); }
#pragma unroll 8
            for (int j = 0; j < 16; j++) {
                const int idx = (FLAGS & 2) ? (base & 16) : base + j;
                float4 A = c_pair[idx]; float2 B = c_w[idx];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    float2 a = __ffma2_rn(make_float2(A.x, A.y), u1[r], __ffma2_rn(make_float2(A.z, A.w), u2[r], nou[r]));
                    if (FLAGS & 1) {
                        fm[r] = __ffma2_rn(a, a, fm[r]);   // keeps 3 FFMA2 per pair, no SHF
                    } else {
                        float2 d = __ffma2_rn(a, a, B);
                        m[r] = __funnelshift_l(__float_as_uint(d.x), m[r], 1);
                        m[r] = __funnelshift_l(__float_as_uint(d.y), m[r], 1);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < R; r++) acc ^= m[r] ^ __float_as_uint(fm[r].x) ^ __float_as_uint(fm[r].y);
        }
#pragma unroll
        for (int r = 0; r < R; r++) nou[r].x += 1e-3f;
    }
    out[blockIdx.x * 128 + threadIdx.x] = acc;
}

__global__ void __launch_bounds__(128) k_peak(unsigned* out, float a, float b, int iters) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = (float) (threadIdx.x + i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = __fmaf_rn(acc[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
    out[blockIdx.x * 128 + threadIdx.x] = __float_as_uint(s);
}
// FFMA2 with per-instruction distinct register operands vs uniform
template <int MODE>
__global__ void __launch_bounds__(128) k_f2(unsigned* out, float a, float b, int iters) {
    float2 acc[8], x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { acc[i] = make_float2(threadIdx.x + i, threadIdx.x - i); x[i] = make_float2(a + i, b - i + threadIdx.x); }
    float2 a2 = make_float2(a, a);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) acc[i] = __ffma2_rn(acc[i], a2, x[i]);           // acc(2 regs) + x(2 regs) reads
            else if (MODE == 1) acc[i] = __ffma2_rn(x[i], x[(i + 3) & 7], acc[i]);   // 6 reg reads
            else if (MODE == 2) acc[i] = __ffma2_rn(acc[i], acc[i], x[i]);  // a*a + B shape
            else if (MODE == 3) { acc[i].x = __fmaf_rn(x[i].x, x[(i + 3) & 7].y, acc[i].x); acc[i].y = __fmaf_rn(x[i].y, x[(i + 5) & 7].x, acc[i].y); } // scalar 3 distinct
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i].x + acc[i].y;
    out[blockIdx.x * 128 + threadIdx.x] = __float_as_uint(s);
}
// SHF throughput
template <int N>
__global__ void __launch_bounds__(128) k_shf(unsigned* out, unsigned a, int iters) {
    unsigned m[N];
#pragma unroll
    for (int i = 0; i < N; i++) m[i] = threadIdx.x * 7 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < N; i++) m[i] = __funnelshift_l(a + i, m[i], 1);
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < N; i++) s ^= m[i];
    out[blockIdx.x * 128 + threadIdx.x] = s;
}

static int g_sms;
static double g_ghz = 1.965;
template <class F> double timeit(F launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    return best * 1e-3;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = g_sms = p.multiProcessorCount;
    std::vector<float4> h(NPRIM);
    for (int i = 0; i < NPRIM; i++) h[i] = make_float4(0.01f * i, 0.2f, -0.03f * i, -(0.04f + i));
    cudaMemcpyToSymbol(c_pair, h.data(), 16 * NPRIM / 2); cudaMemcpyToSymbol(c_w, h.data(), 8 * NPRIM / 2);
    unsigned* out; cudaMalloc(&out, 4 * sms * 16 * 128);
    const int C = 5, iters = 1 << 14;
    // cycles per warp instruction per SMSP, assuming 1.965 GHz
    auto cyc = [&](double sec, double warp_instr_per_warp) { return sec * g_ghz * 1e9 / (warp_instr_per_warp * C /* warps per SMSP */); };
    double t;
    t = timeit([&] { k_peak<<<sms * C, 128>>>(out, 0.999f, 0.001f, iters); });
    printf("FFMA (reuse) : %.3f cyc/instr\n", cyc(t, 16.0 * iters));
    t = timeit([&] { k_f2<0><<<sms * C, 128>>>(out, 0.999f, 0.001f, iters); }); printf("FFMA2 acc*a2+x[i] (4 reg reads): %.3f cyc/instr\n", cyc(t, 8.0 * iters));
    t = timeit([&] { k_f2<1><<<sms * C, 128>>>(out, 0.999f, 0.001f, iters); }); printf("FFMA2 x*y+acc (6 reg reads): %.3f cyc/instr\n", cyc(t, 8.0 * iters));
    t = timeit([&] { k_f2<2><<<sms * C, 128>>>(out, 0.999f, 0.001f, iters); }); printf("FFMA2 acc*acc+x (4 reg reads): %.3f cyc/instr\n", cyc(t, 8.0 * iters));
    t = timeit([&] { k_f2<3><<<sms * C, 128>>>(out, 0.999f, 0.001f, iters); }); printf("FFMA x*y+acc (3 distinct regs): %.3f cyc/instr\n", cyc(t, 16.0 * iters));
    t = timeit([&] { k_shf<8><<<sms * C, 128>>>(out, 12345u, iters); }); printf("SHF x8: %.3f cyc/instr\n", cyc(t, 8.0 * iters));
    t = timeit([&] { k_shf<16><<<sms * C, 128>>>(out, 12345u, iters); }); printf("SHF x16: %.3f cyc/instr\n", cyc(t, 16.0 * iters));
    const int sweeps = 400;
    auto per_test = [&](double sec, int R) { return sec * g_ghz * 1e9 / ((double) C * R * NPRIM * sweeps); };
#define RUN(R, F, name) t = timeit([&] { k_cpair2d<R, F><<<sms * C, 128>>>(out, sweeps); }); printf("%-40s R=%d: %.2f cycles/test\n", name, R, per_test(t, R));
    RUN(2, 0, "full (3 FFMA2 + 2 SHF + LDCU)") RUN(4, 0, "full (3 FFMA2 + 2 SHF + LDCU)")
    RUN(2, 1, "no SHF") RUN(4, 1, "no SHF")
    RUN(2, 2, "no LDCU (same record)") RUN(4, 2, "no LDCU (same record)")
    RUN(2, 3, "no SHF, no LDCU") RUN(4, 3, "no SHF, no LDCU")
    return 0;
}
