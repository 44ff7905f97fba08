// sdrplusplus_b200/csrc/chanpfb.cuh -- BASELINE config 3: M-channel critically sampled polyphase filter-bank channelizer
// (M = 256, P = 127 taps per branch; prototype = the reference's taps::windowedSinc<float>(M*P, fs/(2M), fs, nuttall),
// core/src/dsp/taps/windowed_sinc.h:31-34).  NOT a reference feature: the definition is the
// direct form   y_k[m] = sum_t h[t] x[n0 + t] e^{-j 2 pi k (n0 + t) / M},  n0 = m M + (M - 1) - (T - 1),  T = M P
// (translate by -k fs/M, T-tap FIR in the reference's correlation form, keep the samples n = m M + M - 1), which the
// float64 oracle in tests/test_gpu_channelizer.py evaluates literally.  With t = p M + r the phasor depends on r only:
//     u_r[m] = sum_p h[p M + r] x[(m + p) M + r - (P - 1) M]            (k_chan_branch: real taps, shared by all channels)
//     y_k[m] = sum_r u_r[m] e^{-j 2 pi k r / M}                          (k_chan_fft: one M-point DFT per output time)
// The branch outputs of a chunk (8 B per input sample) stay in L2 between the two kernels.
#pragma once

#define CH_MT 256            // output times per CTA of the branch kernel
#define CH_R 8               // consecutive output times per thread


// grid (M / 32, ceil(n_out / CH_MT)); 256 threads: lane = branch inside the group, warp = 32 output times
__global__ void __launch_bounds__(256, 2) k_chan_branch(const __grid_constant__ ChanParams p) {
    extern __shared__ __align__(16) float2 ch_sm[];
    const int P = p.P, M = p.M;
    float* hs = reinterpret_cast<float*>(ch_sm);                     // [P][32]
    float2* xs = ch_sm + (P * 32 + 1) / 2;                           // [(CH_MT + P - 1)][32]
    const int g = blockIdx.x, m0 = blockIdx.y * CH_MT;
    const int rows = min(CH_MT, p.n_out - m0) + P - 1;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int i = tid; i < P * 32; i += 256) { hs[i] = __ldg(p.h + (size_t)g * P * 32 + i); }
    // rows of 32 samples (256 B), 2 KB apart in the stream: 16-byte copies, 16 per row
    for (int i = tid; i < rows * 16; i += 256) {
        const int row = i >> 4, c = i & 15;
        const float4* src = reinterpret_cast<const float4*>(p.in + (size_t)(m0 + row) * M + g * 32) + c;
        unsigned dst = (unsigned)__cvta_generic_to_shared(reinterpret_cast<float4*>(xs + row * 32) + c);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    // warp w: output times m0 + 32 w .. + 31, CH_R at a time; lane = branch r
    for (int q = 0; q < 32 / CH_R; q++) {
        const int ml = w * 32 + q * CH_R;                            // local first output time
        if (m0 + ml >= p.n_out) { break; }
        float2 acc[CH_R], win[CH_R];
#pragma unroll
        for (int i = 0; i < CH_R; i++) { acc[i] = make_float2(0.f, 0.f); win[i] = xs[(ml + i) * 32 + lane]; }
        int pp = 0;
        for (; pp + CH_R <= P; pp += CH_R) {
#pragma unroll
            for (int k = 0; k < CH_R; k++) {
                const float hv = hs[(pp + k) * 32 + lane];
#pragma unroll
                for (int i = 0; i < CH_R; i++) { acc[i] = ffma2(make_float2(hv, hv), win[(k + i) % CH_R], acc[i]); }
                win[k] = xs[(ml + pp + k + CH_R) * 32 + lane];       // slide: rows up to ml + P + CH_R - 2 < rows + CH_R (padded)
            }
        }
        for (; pp < P; pp++) {                                        // P % CH_R leftover taps
            const float hv = hs[pp * 32 + lane];
            const int k = pp % CH_R;
#pragma unroll
            for (int i = 0; i < CH_R; i++) {
                // win[(k + i) % R] holds row ml + pp + i as long as the slide above kept running; past it, read directly
                acc[i] = ffma2(make_float2(hv, hv), xs[(ml + pp + i) * 32 + lane], acc[i]);
            }
            (void)k;
        }
#pragma unroll
        for (int i = 0; i < CH_R; i++) {
            const int m = m0 + ml + i;
            if (m < p.n_out) { p.u[(size_t)m * M + g * 32 + lane] = acc[i]; }
        }
    }
}

// one CTA = 8 output times: M-point forward DFT across the branches (M = 256), natural-order output
__global__ void __launch_bounds__(256) k_chan_fft(const float2* __restrict__ u, float2* __restrict__ y, const float2* __restrict__ tw, int n_out) {
    extern __shared__ __align__(16) float2 ch_fs[];
    constexpr int M = 256, LOGM = 8, RR = 8;
    const int pitch = M + 1;
    const int m0 = blockIdx.x * RR;
    for (int t = threadIdx.x; t < RR * M; t += blockDim.x) {
        const int rl = t >> LOGM, r = t & (M - 1);
        ch_fs[padf(rl * pitch + r)] = (m0 + rl < n_out) ? __ldg(u + (size_t)(m0 + rl) * M + r) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    fft_dif_smem<false>(ch_fs, LOGM, RR, pitch, tw, LOGM);
    for (int t = threadIdx.x; t < RR * M; t += blockDim.x) {
        const int rl = t >> LOGM, k = t & (M - 1);
        if (m0 + rl < n_out) { y[(size_t)(m0 + rl) * M + k] = ch_fs[padf(rl * pitch + bitrev(k, LOGM))]; }
    }
}

cudaError_t launch_channelizer(const ChanParams& p, float2* y, const float2* tw, cudaStream_t s, int* nlaunch) {
    if (p.n_out <= 0) { return cudaSuccess; }
    if (p.M != 256 || (p.M & 31)) { return cudaErrorInvalidValue; }
    const size_t smem1 = ((size_t)(p.P * 32 + 1) / 2 + (size_t)(CH_MT + p.P - 1 + CH_R) * 32) * sizeof(float2);
    if ((int)smem1 > kernels_max_smem_optin()) { return cudaErrorInvalidValue; }
    cudaError_t e = set_smem(k_chan_branch, smem1);
    if (e != cudaSuccess) { return e; }
    dim3 grid((unsigned)(p.M / 32), (unsigned)((p.n_out + CH_MT - 1) / CH_MT));
    k_chan_branch<<<grid, 256, smem1, s>>>(p);
    const size_t smem2 = ((size_t)8 * 257 + (8 * 257 >> 4) + 2) * sizeof(float2);
    k_chan_fft<<<(unsigned)((p.n_out + 7) / 8), 256, smem2, s>>>(p.u, y, tw, p.n_out);
    if (nlaunch) { *nlaunch += 2; }
    return cudaGetLastError();
}
