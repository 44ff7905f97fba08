// tools/ubench.cu -- micro-benchmarks that size the stage-1 kernel design on the actual B200:
//   fp32 FFMA vs packed FFMA2 issue rate, and the FFMA2 + uniform/shared tap-load mixes the tile kernel uses.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench tools/ubench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}

template <int NACC>
__global__ void k_ffma(float* out, int iters, float a, float b) {
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) { acc[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) { acc[i] = fmaf(acc[i], a, b); }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) { s += acc[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_ffma2(float2* out, int iters, float a, float b) {
    float2 acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) { acc[i] = make_float2(threadIdx.x + i, i); }
    const float2 bb = make_float2(b, b);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) { acc[i] = ffma2(acc[i], make_float2(a, a), bb); }
    }
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < NACC; i++) { s.x += acc[i].x; s.y += acc[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FFMA2 with the multiplier coming from shared memory (broadcast LDS.64 per 2*NX FFMA2) -- the tile kernel's mix
template <int NX>
__global__ void k_mix_lds(float2* out, int iters) {
    __shared__ float2 taps[256];
    __shared__ float2 xs[1024 + 64];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { taps[i] = make_float2(1.0f + i * 1e-6f, 1e-6f * i); }
    for (int i = threadIdx.x; i < 1024 + 64; i += blockDim.x) { xs[i] = make_float2(i * 1e-3f, 1.0f); }
    __syncthreads();
    float2 A[NX], B[NX];
#pragma unroll
    for (int i = 0; i < NX; i++) { A[i] = make_float2(0, 0); B[i] = make_float2(0, 0); }
    for (int it = 0; it < iters; it++) {
        float2 x[NX];
#pragma unroll
        for (int i = 0; i < NX; i++) { x[i] = xs[(threadIdx.x + 32 * i + it) & 1023]; }
#pragma unroll
        for (int t = 0; t < 16; t++) {
            float2 g = taps[(it * 16 + t) & 255];
#pragma unroll
            for (int i = 0; i < NX; i++) {
                A[i] = ffma2(make_float2(g.x, g.x), x[i], A[i]);
                B[i] = ffma2(make_float2(g.y, g.y), x[i], B[i]);
            }
        }
    }
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < NX; i++) { s.x += A[i].x - B[i].y; s.y += A[i].y + B[i].x; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float time_ms(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { printf("no device\n"); return 1; }
    printf("device %s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    const int sms = p.multiProcessorCount;
    float2* out;
    cudaMalloc(&out, sizeof(float2) * 1024 * 1024 * 4);
    const int iters = 4096;
    for (int warps = 4; warps <= 32; warps *= 2) {
        int thr = warps * 32 > 1024 ? 1024 : warps * 32;
        int blocks = sms * (warps * 32 / thr);
        float ms = time_ms([&] { k_ffma<16><<<blocks, thr>>>((float*)out, iters, 1.0001f, 0.5f); });
        double fma = (double)blocks * thr * iters * 16;
        printf("FFMA   x16acc  %2d warps/SM : %7.3f ms  %.2f TFMA/s  (%.1f FMA/clk/SM @%.0f MHz nominal)\n", warps, ms, fma / ms * 1e-9,
               fma / ms * 1e-3 / sms / (p.clockRate * 1e3) * 1e6 / 1e3 * 1e3, p.clockRate * 1e-3);
        ms = time_ms([&] { k_ffma2<16><<<blocks, thr>>>(out, iters, 1.0001f, 0.5f); });
        fma = (double)blocks * thr * iters * 16 * 2;
        printf("FFMA2  x16acc  %2d warps/SM : %7.3f ms  %.2f TFMA/s\n", warps, ms, fma / ms * 1e-9);
    }
    for (int warps = 4; warps <= 16; warps *= 2) {
        int thr = warps * 32;
        float ms = time_ms([&] { k_mix_lds<4><<<sms, thr>>>(out, iters); });
        double fma = (double)sms * thr * iters * 16 * 4 * 2 * 2;
        printf("MIX LDS NX=4   %2d warps/SM : %7.3f ms  %.2f TFMA/s\n", warps, ms, fma / ms * 1e-9);
        ms = time_ms([&] { k_mix_lds<8><<<sms, thr>>>(out, iters); });
        fma = (double)sms * thr * iters * 16 * 8 * 2 * 2;
        printf("MIX LDS NX=8   %2d warps/SM : %7.3f ms  %.2f TFMA/s\n", warps, ms, fma / ms * 1e-9);
    }
    return 0;
}
