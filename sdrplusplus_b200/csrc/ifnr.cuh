// sdrplusplus_b200/csrc/ifnr.cuh -- FM IF noise reduction of the radio module's IF chain (included by kernels.cu):
//   k_fmif_pow2<N>, k_fmif_direct   noise_reduction::FMIF::process (core/src/dsp/noise_reduction/fm_if.h:44-77): for EVERY
//       output sample a Nuttall-windowed `bins`-point transform of the last `bins` input samples, the strongest bin alone
//       transformed back, element bins/2 of that kept.  The reference does this sample by sample with two FFTW plans; here a
//       thread owns an output sample and the whole transform lives in its registers.
// Which bin is the strongest decides the output, so the forward transform follows the oracle's leaf layer operation for
// operation (oracle/offt.h: radix-4 / radix-2 Stockham passes for a power of two, the defining sum in ascending order for the
// 9 / 15 / 31-bin presets of radio_module.h:31-36), with explicitly rounded multiplies and adds: no FMA contraction.
// The noise blanker of the same chain (noise_blanker.h:38-57) is a sequential recurrence: k_seq kind 3 (kernels.cu).
#pragma once

#define FMIF_THREADS 128
#define FMIF_MAXBINS 64

__device__ __forceinline__ float2 fm_cmul(float2 a, float2 b) {          // offt_mul: four products, one subtract, one add
    return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float fm_mag(float2 a) { return __fsqrt_rn(__fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y))); }

// one pass of offt_forward on register arrays: x -> y (radix 4 while n >= 4, one radix-2 pass at the end), result in `out`
template <int N, int n, int s>
__device__ __forceinline__ void fm_pass(float2 (&x)[N], float2 (&y)[N], float2 (&out)[N], const float2* __restrict__ tw) {
    if constexpr (n >= 4) {
        constexpr int n1 = n / 4, n2 = n / 2, n3 = n1 + n2, tstep = N / n;
#pragma unroll
        for (int p = 0; p < n1; p++) {
            const float2 w1 = tw[p * tstep], w2 = tw[2 * p * tstep], w3 = tw[3 * p * tstep];
#pragma unroll
            for (int q = 0; q < s; q++) {
                const float2 a = x[q + s * p], b = x[q + s * (p + n1)], c = x[q + s * (p + n2)], d = x[q + s * (p + n3)];
                const float2 apc = make_float2(__fadd_rn(a.x, c.x), __fadd_rn(a.y, c.y));
                const float2 amc = make_float2(__fsub_rn(a.x, c.x), __fsub_rn(a.y, c.y));
                const float2 bpd = make_float2(__fadd_rn(b.x, d.x), __fadd_rn(b.y, d.y));
                const float2 jbmd = make_float2(-__fsub_rn(b.y, d.y), __fsub_rn(b.x, d.x));        // j (b - d)
                const float2 t0 = make_float2(__fadd_rn(apc.x, bpd.x), __fadd_rn(apc.y, bpd.y));
                const float2 t1 = make_float2(__fsub_rn(amc.x, jbmd.x), __fsub_rn(amc.y, jbmd.y));
                const float2 t2 = make_float2(__fsub_rn(apc.x, bpd.x), __fsub_rn(apc.y, bpd.y));
                const float2 t3 = make_float2(__fadd_rn(amc.x, jbmd.x), __fadd_rn(amc.y, jbmd.y));
                y[q + s * (4 * p + 0)] = t0;
                y[q + s * (4 * p + 1)] = fm_cmul(t1, w1);
                y[q + s * (4 * p + 2)] = fm_cmul(t2, w2);
                y[q + s * (4 * p + 3)] = fm_cmul(t3, w3);
            }
        }
        fm_pass<N, n / 4, s * 4>(y, x, out, tw);
    }
    else if constexpr (n == 2) {
#pragma unroll
        for (int q = 0; q < s; q++) {
            const float2 a = x[q], b = x[q + s];
            out[q] = make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
            out[q + s] = make_float2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y));
        }
    }
    else {
#pragma unroll
        for (int q = 0; q < N; q++) { out[q] = x[q]; }
    }
}

// shared staging common to both kernels: the CTA's input tile, the window and the twiddles
struct FmTile {
    float2 x[FMIF_THREADS + FMIF_MAXBINS];
    float win[FMIF_MAXBINS];
    float2 tw[FMIF_MAXBINS];
};
__device__ __forceinline__ void fm_fill(FmTile& T, const FmIfJob& J, int i0) {
    const int N = J.bins;
    for (int k = threadIdx.x; k < FMIF_THREADS + N - 1; k += FMIF_THREADS) {
        T.x[k] = (i0 + k < J.n + N - 1) ? __ldg(J.in + i0 + k) : make_float2(0.0f, 0.0f);
    }
    if (threadIdx.x < N) {
        T.win[threadIdx.x] = __ldg(J.win + threadIdx.x);
        T.tw[threadIdx.x] = __ldg(J.tw + threadIdx.x);
    }
    __syncthreads();
}
// the single non-zero bin transformed back (offt_backward: defining sum, the zero bins add exact zeros), element N / 2
__device__ __forceinline__ float2 fm_back(float2 X, int idx, int N, const float2* tw) {
    const float2 w = tw[(idx * (N / 2)) % N];
    const float2 pr = fm_cmul(X, make_float2(w.x, -w.y));
    return make_float2(__fadd_rn(0.0f, pr.x), __fadd_rn(0.0f, pr.y));
}

template <int N>
__global__ void __launch_bounds__(FMIF_THREADS) k_fmif_pow2(const __grid_constant__ FmIfParams p) {
    __shared__ FmTile T;
    const FmIfJob& J = p.job[blockIdx.y];
    const int i0 = blockIdx.x * FMIF_THREADS;
    if (i0 >= J.n) { return; }
    fm_fill(T, J, i0);
    const int i = i0 + threadIdx.x;
    if (i >= J.n) { return; }
    float2 a[N], b[N], X[N];
#pragma unroll
    for (int j = 0; j < N; j++) {                                       // volk_32fc_32f_multiply_32fc
        const float2 v = T.x[threadIdx.x + j];
        const float w = T.win[j];
        a[j] = make_float2(__fmul_rn(v.x, w), __fmul_rn(v.y, w));
    }
    fm_pass<N, N, 1>(a, b, X, T.tw);
    // volk_32fc_magnitude_32f + volk_32f_index_max_32u: first index of the largest magnitude
    float best = fm_mag(X[0]);
    float2 bx = X[0];
    int idx = 0;
#pragma unroll
    for (int k = 1; k < N; k++) {
        const float m = fm_mag(X[k]);
        if (m > best) { best = m; bx = X[k]; idx = k; }
    }
    J.out[i] = fm_back(bx, idx, N, T.tw);
}

// any bin count: the defining sum, j ascending (offt_forward's path for a size that is not a power of two)
__global__ void __launch_bounds__(FMIF_THREADS) k_fmif_direct(const __grid_constant__ FmIfParams p) {
    __shared__ FmTile T;
    const FmIfJob& J = p.job[blockIdx.y];
    const int i0 = blockIdx.x * FMIF_THREADS;
    if (i0 >= J.n) { return; }
    fm_fill(T, J, i0);
    const int i = i0 + threadIdx.x;
    if (i >= J.n) { return; }
    const int N = J.bins;
    float best = -1.0f;
    float2 bx = make_float2(0.0f, 0.0f);
    int idx = 0;
    for (int k = 0; k < N; k++) {
        float2 acc = make_float2(0.0f, 0.0f);
        int r = 0;                                                       // (k j) mod N
        for (int j = 0; j < N; j++) {
            const float2 v = T.x[threadIdx.x + j];
            const float w = T.win[j];
            const float2 pr = fm_cmul(make_float2(__fmul_rn(v.x, w), __fmul_rn(v.y, w)), T.tw[r]);
            acc = make_float2(__fadd_rn(acc.x, pr.x), __fadd_rn(acc.y, pr.y));
            r += k;
            if (r >= N) { r -= N; }
        }
        const float m = fm_mag(acc);
        if (k == 0 || m > best) { best = m; bx = acc; idx = k; }
    }
    J.out[i] = fm_back(bx, idx, N, T.tw);
}

bool fmif_supported(int bins) { return bins >= 2 && bins <= FMIF_MAXBINS; }
// every job of p: the same bin count
cudaError_t launch_fmif(const FmIfParams& p, cudaStream_t s) {
    if (p.njobs <= 0 || p.max_n <= 0) { return cudaSuccess; }
    const int N = p.job[0].bins;
    dim3 grid((unsigned)((p.max_n + FMIF_THREADS - 1) / FMIF_THREADS), (unsigned)p.njobs);
    switch (N) {
    case 2: k_fmif_pow2<2><<<grid, FMIF_THREADS, 0, s>>>(p); break;
    case 4: k_fmif_pow2<4><<<grid, FMIF_THREADS, 0, s>>>(p); break;
    case 8: k_fmif_pow2<8><<<grid, FMIF_THREADS, 0, s>>>(p); break;
    case 16: k_fmif_pow2<16><<<grid, FMIF_THREADS, 0, s>>>(p); break;
    case 32: k_fmif_pow2<32><<<grid, FMIF_THREADS, 0, s>>>(p); break;
    case 64: k_fmif_pow2<64><<<grid, FMIF_THREADS, 0, s>>>(p); break;
    default:
        if (!fmif_supported(N)) { return cudaErrorInvalidValue; }
        k_fmif_direct<<<grid, FMIF_THREADS, 0, s>>>(p);
        break;
    }
    return cudaGetLastError();
}
