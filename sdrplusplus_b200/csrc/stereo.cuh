// sdrplusplus_b200/csrc/stereo.cuh -- the stereo branch of demod::BroadcastFM behind the discriminator
// (core/src/dsp/demod/broadcast_fm.h:147-190), at the demodulator's IF rate (250 kS/s).  Included by kernels.cu.
//
//   m = discriminator output (L+R multiplex)            [hist | data] with hist = pilot taps - 1
//   k_st_pilot  p[i]   = sum_k taps[k] * (m[i+k], 0)     RealToComplex + FIR<complex_t,complex_t> with
//                                                         taps::bandPass<complex_t>(18750, 19250, 3000, fs, odd)   (:44-46,150-153)
//   k_st_pll    vco[i] = phasor(phase);  advance(normalizePhase(arg p[i] - phase))        loop::PLL (pll.h:64-70) over
//                                                         PhaseControlLoop (phase_control_loop.h:58-85): ONE thread per VFO,
//                                                         the recurrence is the reference's own fp32 statements
//   k_st_mix    z = (m[i - delay], 0) * conj(vco) * conj(vco);  lmr = 2 Re z;  l = m_d + lmr,  r = m_d - lmr      (:156-177)
// The two audio low-passes + LRToStereo behind it are one FIR over (l, r) pairs (the complex-data / real-tap kernels).
#pragma once


__global__ void __launch_bounds__(256) k_st_pilot(const __grid_constant__ StParams p) {
    const StJob& J = p.job[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= J.n) { return; }
    const float* __restrict__ x = J.in + i;
    float ar = 0.0f, ai = 0.0f;
    for (int k = 0; k < J.ntaps; k++) {
        const float v = __ldg(x + k);
        const float2 t = __ldg(J.taps + k);
        ar = fmaf(v, t.x, ar);
        ai = fmaf(v, t.y, ai);
    }
    J.p[i] = make_float2(ar, ai);
}

// one thread per VFO; explicit roundings: the loop filter is a feedback loop, its decisions follow the reference's fp32 steps
__global__ void k_st_pll(const __grid_constant__ StParams p) {
    const StJob& J = p.job[blockIdx.x];
    if (threadIdx.x != 0) { return; }
    float phase = J.state[0], freq = J.state[1];
    const float pi = FL_M_PI_REF, two_pi = __fsub_rn(FL_M_PI_REF, -FL_M_PI_REF);     // phaseDelta = maxPhase - minPhase
    for (int i = 0; i < J.n; i++) {
        float sn, cs;
        sincosf(phase, &sn, &cs);
        J.vco[i] = make_float2(cs, sn);
        const float2 v = J.p[i];
        float err = __fsub_rn(atan2f(v.y, v.x), phase);
        if (err > pi) { err = __fsub_rn(err, 2.0f * pi); }
        else if (err <= -pi) { err = __fadd_rn(err, 2.0f * pi); }
        freq = __fadd_rn(freq, __fmul_rn(J.beta, err));
        if (freq > J.max_freq) { freq = J.max_freq; }
        else if (freq < J.min_freq) { freq = J.min_freq; }
        phase = __fadd_rn(phase, __fadd_rn(freq, __fmul_rn(J.alpha, err)));
        while (phase > pi) { phase = __fsub_rn(phase, two_pi); }
        while (phase < -pi) { phase = __fadd_rn(phase, two_pi); }
    }
    J.state[0] = phase;
    J.state[1] = freq;
}

__global__ void __launch_bounds__(256) k_st_mix(const __grid_constant__ StParams p) {
    const StJob& J = p.job[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= J.n) { return; }
    // math::Delay by `delay` samples: sample i - delay of the data = index hist + i - delay of [hist | data]
    const float md = __ldg(J.in + (J.ntaps - 1) + i - J.delay);
    const float2 v = J.vco[i];
    const float cr = v.x, ci = -v.y;                                        // math::Conjugate
    // (md, 0) * c  then  * c again, volk_32fc_x2_multiply_32fc order (products, then the difference / sum)
    const float t1r = __fsub_rn(__fmul_rn(md, cr), __fmul_rn(0.0f, ci));
    const float t1i = __fadd_rn(__fmul_rn(md, ci), __fmul_rn(0.0f, cr));
    const float t2r = __fsub_rn(__fmul_rn(t1r, cr), __fmul_rn(t1i, ci));
    const float lmr = __fmul_rn(t2r, 2.0f);
    J.out[i] = make_float2(__fadd_rn(md, lmr), __fsub_rn(md, lmr));
}

cudaError_t launch_stereo(const StParams& p, cudaStream_t s, int* nlaunch) {
    if (p.njobs <= 0 || p.max_n <= 0) { return cudaSuccess; }
    dim3 grid((unsigned)((p.max_n + 255) / 256), (unsigned)p.njobs);
    k_st_pilot<<<grid, 256, 0, s>>>(p);
    k_st_pll<<<p.njobs, 32, 0, s>>>(p);
    k_st_mix<<<grid, 256, 0, s>>>(p);
    if (nlaunch) { *nlaunch += 3; }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// noise_reduction::PowerSquelch (core/src/dsp/noise_reduction/power_squelch.h:33-50): the mean amplitude of the CHUNK decides
// whether the chunk passes or is zeroed: 10 log10(mean |x|) >= level.  k_sq_sum: per-CTA partial sums of |x| (fp32, then
// combined in double so that the decision does not depend on the grid); k_sq_gate: every CTA reduces the partials, applies the
// reference's float expression and copies or clears its slice.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sq_sum(const __grid_constant__ SqParams p) {
    const SqJob& J = p.job[blockIdx.y];
    __shared__ float ws[8];
    float acc = 0.0f;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < J.n; i += gridDim.x * 256) {
        const float2 v = __ldg(J.in + i);
        acc += __fsqrt_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)));
    }
    for (int d = 16; d > 0; d >>= 1) { acc += __shfl_down_sync(0xffffffffu, acc, d); }
    if ((threadIdx.x & 31) == 0) { ws[threadIdx.x >> 5] = acc; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int k = 0; k < 8; k++) { t += ws[k]; }
        J.partial[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(256) k_sq_gate(const __grid_constant__ SqParams p, int nparts) {
    const SqJob& J = p.job[blockIdx.y];
    double tot = 0.0;
    for (int k = 0; k < nparts; k++) { tot += (double)J.partial[k]; }
    const float mean = (float)tot / (float)J.n;
    const bool open = 10.0f * log10f(mean) >= J.level;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < J.n; i += gridDim.x * 256) {
        J.out[i] = open ? __ldg(J.in + i) : make_float2(0.0f, 0.0f);
    }
}
cudaError_t launch_squelch(const SqParams& p, cudaStream_t s, int* nlaunch) {
    if (p.njobs <= 0 || p.max_n <= 0) { return cudaSuccess; }
    int nparts = (p.max_n + 256 * 16 - 1) / (256 * 16);
    if (nparts > SQ_MAXPARTS) { nparts = SQ_MAXPARTS; }
    if (nparts < 1) { nparts = 1; }
    dim3 grid((unsigned)nparts, (unsigned)p.njobs);
    k_sq_sum<<<grid, 256, 0, s>>>(p);
    k_sq_gate<<<grid, 256, 0, s>>>(p, nparts);
    if (nlaunch) { *nlaunch += 2; }
    return cudaGetLastError();
}
