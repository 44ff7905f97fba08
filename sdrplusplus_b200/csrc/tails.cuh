// sdrplusplus_b200/csrc/tails.cuh -- shared-memory tiled versions of the post-stage-1 kernels (included by
// kernels.cu).  The v0 kernels in kernels.cu (one thread per output, operands through L1) stay as the fallback
// for shapes these tiles do not cover (very long filters).
//
// Common scheme: a CTA stages the input span of its outputs in shared memory with a PADDED layout
// (one pad slot every `padq` samples) so that lanes whose windows start `padq` samples apart hit different
// banks; taps sit in shared memory and are read as warp-wide broadcasts.
#pragma once

#define TAIL_MAX_TAPS 1024

__device__ __forceinline__ int pad_idx(int n, int padq_log2) { return n + (n >> padq_log2); }

// ---- complex data x real taps, decimation 1, R consecutive outputs per thread with a sliding register window
//      (RxVFO channel filter: FIR<complex_t,float>, fir.h:62-83) ----
#define FC2_R 8
#define FC2_THREADS 64
__global__ void __launch_bounds__(FC2_THREADS) k_fir_c2(const __grid_constant__ FirParams p) {
    extern __shared__ __align__(16) float2 smem[];
    const FirJob& J = p.job[blockIdx.y];
    const int OB = FC2_THREADS * FC2_R;
    const int m0 = blockIdx.x * OB;
    if (m0 >= J.n_out) { return; }
    const int T = J.ntaps;
    const int nout = min(OB, J.n_out - m0);
    const int span = nout + T - 1;                       // input samples needed (decim 1, offset 0)
    float* taps = reinterpret_cast<float*>(smem);        // [T] (rounded up to even)
    float2* X = smem + ((T + 1) >> 1);                   // padded every 8
    for (int k = threadIdx.x; k < T; k += blockDim.x) { taps[k] = __ldg(J.taps + k); }
    const float2* __restrict__ src = J.in + (size_t)J.offset + m0;
    for (int n = threadIdx.x; n < span + FC2_R; n += blockDim.x) {
        X[pad_idx(n, 3)] = (n < span) ? __ldg(src + n) : make_float2(0.0f, 0.0f);
    }
    __syncthreads();
    const int mo = threadIdx.x * FC2_R;
    if (mo >= nout) { return; }
    float2 acc[FC2_R], w[FC2_R];
#pragma unroll
    for (int i = 0; i < FC2_R; i++) { acc[i] = make_float2(0.f, 0.f); w[i] = X[pad_idx(mo + i, 3)]; }
    for (int k0 = 0; k0 < T; k0 += FC2_R) {
#pragma unroll
        for (int kk = 0; kk < FC2_R; kk++) {
            const int k = k0 + kk;
            if (k < T) {
                const float h = taps[k];
#pragma unroll
                for (int i = 0; i < FC2_R; i++) { acc[i] = ffma2(make_float2(h, h), w[(kk + i) % FC2_R], acc[i]); }
                // slide: the slot that held x[mo+k] is no longer needed, refill it with x[mo+k+R]
                w[kk % FC2_R] = X[pad_idx(mo + k + FC2_R, 3)];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < FC2_R; i++) {
        if (mo + i < nout) { J.out[m0 + mo + i] = acc[i]; }
    }
}

// ---- complex data x real taps, decimation D > 1, one output per thread (DecimatingFIR stages 2..k) ----
#define FCD_THREADS 256
__global__ void __launch_bounds__(FCD_THREADS) k_fir_cd(const __grid_constant__ FirParams p, int padq_log2) {
    extern __shared__ __align__(16) float2 smem[];
    const FirJob& J = p.job[blockIdx.y];
    const int m0 = blockIdx.x * FCD_THREADS;
    if (m0 >= J.n_out) { return; }
    const int T = J.ntaps, D = J.decim;
    const int nout = min(FCD_THREADS, J.n_out - m0);
    const int span = (nout - 1) * D + T;
    float* taps = reinterpret_cast<float*>(smem);
    float2* X = smem + ((T + 1) >> 1);
    for (int k = threadIdx.x; k < T; k += blockDim.x) { taps[k] = __ldg(J.taps + k); }
    const float2* __restrict__ src = J.in + (size_t)J.offset + (size_t)m0 * D;
    for (int n = threadIdx.x; n < span; n += blockDim.x) { X[pad_idx(n, padq_log2)] = __ldg(src + n); }
    __syncthreads();
    if ((int)threadIdx.x >= nout) { return; }
    const int base = threadIdx.x * D;
    float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
    int k = 0;
    for (; k + 1 < T; k += 2) {
        const float h0 = taps[k], h1 = taps[k + 1];
        a0 = ffma2(make_float2(h0, h0), X[pad_idx(base + k, padq_log2)], a0);
        a1 = ffma2(make_float2(h1, h1), X[pad_idx(base + k + 1, padq_log2)], a1);
    }
    if (k < T) {
        const float h0 = taps[k];
        a0 = ffma2(make_float2(h0, h0), X[pad_idx(base + k, padq_log2)], a0);
    }
    J.out[m0 + threadIdx.x] = make_float2(a0.x + a1.x, a0.y + a1.y);
}

// ---- real data x real taps, R consecutive outputs per thread, optional stereo duplication on store
//      (audio low-pass: FIR<float,float> + LRToStereo, fir.h:69, l_r_to_stereo.h:21) ----
#define FR2_R 8
#define FR2_THREADS 64
__global__ void __launch_bounds__(FR2_THREADS) k_fir_r2(const __grid_constant__ FirRParams p) {
    extern __shared__ __align__(16) float smemf[];
    const FirRJob& J = p.job[blockIdx.y];
    const int OB = FR2_THREADS * FR2_R;
    const int m0 = blockIdx.x * OB;
    if (m0 >= J.n_out) { return; }
    const int T = J.ntaps;
    const int nout = min(OB, J.n_out - m0);
    const int span = nout + T - 1;
    float* taps = smemf;
    float* X = smemf + T;                                // padded every 8
    for (int k = threadIdx.x; k < T; k += blockDim.x) { taps[k] = __ldg(J.taps + k); }
    const float* __restrict__ src = J.in + m0;
    for (int n = threadIdx.x; n < span + FR2_R; n += blockDim.x) { X[pad_idx(n, 3)] = (n < span) ? __ldg(src + n) : 0.0f; }
    __syncthreads();
    const int mo = threadIdx.x * FR2_R;
    if (mo >= nout) { return; }
    float acc[FR2_R], w[FR2_R];
#pragma unroll
    for (int i = 0; i < FR2_R; i++) { acc[i] = 0.0f; w[i] = X[pad_idx(mo + i, 3)]; }
    for (int k0 = 0; k0 < T; k0 += FR2_R) {
#pragma unroll
        for (int kk = 0; kk < FR2_R; kk++) {
            const int k = k0 + kk;
            if (k < T) {
                const float h = taps[k];
#pragma unroll
                for (int i = 0; i < FR2_R; i++) { acc[i] = fmaf(h, w[(kk + i) % FR2_R], acc[i]); }
                w[kk % FR2_R] = X[pad_idx(mo + k + FR2_R, 3)];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < FR2_R; i++) {
        if (mo + i < nout) {
            if (J.stereo) { reinterpret_cast<float2*>(J.out)[m0 + mo + i] = make_float2(acc[i], acc[i]); }
            else { J.out[m0 + mo + i] = acc[i]; }
        }
    }
}

// ---- polyphase rational resampler: input span and (when it fits) the whole bank in shared memory ----
#define PL2_THREADS 256
__global__ void __launch_bounds__(PL2_THREADS) k_poly2(const __grid_constant__ PolyParams p, int span_cap, int bank_in_smem) {
    extern __shared__ __align__(16) float2 smem[];
    const PolyJob& J = p.job[blockIdx.y];
    const int m0 = blockIdx.x * PL2_THREADS;
    if (m0 >= J.n_out) { return; }
    const int nout = min(PL2_THREADS, J.n_out - m0);
    const int tpp = J.tpp, L = J.interp;
    // first / last input index touched by this CTA's outputs
    const long long t_first = (long long)J.phase0 + (long long)m0 * J.decim;
    const long long t_last = (long long)J.phase0 + (long long)(m0 + nout - 1) * J.decim;
    const long long off_first = (long long)J.offset0 + t_first / L;
    const int span = (int)((long long)J.offset0 + t_last / L - off_first) + tpp;
    float2* X = smem;                                                    // [span_cap]
    float* bank = reinterpret_cast<float*>(smem + span_cap);             // [L][tpp | 1] when it fits
    const int pitch = tpp | 1;
    for (int n = threadIdx.x; n < span; n += blockDim.x) { X[n] = __ldg(J.in + off_first + n); }
    if (bank_in_smem) {
        for (int idx = threadIdx.x; idx < L * tpp; idx += blockDim.x) {
            int ph = idx / tpp, k = idx - ph * tpp;
            bank[ph * pitch + k] = __ldg(J.bank + idx);
        }
    }
    __syncthreads();
    if ((int)threadIdx.x >= nout) { return; }
    const long long t = t_first + (long long)threadIdx.x * J.decim;
    const int xo = (int)((long long)J.offset0 + t / L - off_first);
    const int ph = (int)(t % L);
    float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
    if (bank_in_smem) {
        const float* h = bank + ph * pitch;
        int k = 0;
        for (; k + 1 < tpp; k += 2) {
            a0 = ffma2(make_float2(h[k], h[k]), X[xo + k], a0);
            a1 = ffma2(make_float2(h[k + 1], h[k + 1]), X[xo + k + 1], a1);
        }
        if (k < tpp) { a0 = ffma2(make_float2(h[k], h[k]), X[xo + k], a0); }
    }
    else {
        const float* __restrict__ h = J.bank + (size_t)ph * tpp;
        for (int k = 0; k < tpp; k++) {
            const float c = __ldg(h + k);
            a0 = ffma2(make_float2(c, c), X[xo + k], a0);
        }
    }
    J.out[m0 + threadIdx.x] = make_float2(a0.x + a1.x, a0.y + a1.y);
}
