// sdrplusplus_b200/csrc/preproc.cuh -- IQFrontEnd's pre-processing chain at the full input rate (included by kernels.cu):
//   correction::DCBlocker<complex_t>   y[i] = x[i] - o;  o += y[i] * rate        (core/src/dsp/correction/dc_blocker.h:54-60,
//                                       rate = 50 / fs: core/src/signal_path/iq_frontend.h:55-57)
//   math::Conjugate                     y = conj(x)                                (core/src/dsp/math/conjugate.h:13,
//                                       IQFrontEnd::setInvertIQ, iq_frontend.cpp:117-119)
// in the reference's order (iq_frontend.cpp:36-39: decimator -> DC blocker -> conjugate).  The decimator in front of them
// is the PowerDecimator chain the library already has (Chain::add_power_decim).
//
// The blocker is a first-order linear recurrence over the whole stream: o[i+1] = (1 - r) o[i] + r x[i].  Three passes:
//   1. k_dcb_reduce  every segment of DCB_SEG samples -> its affine map o_out = (1 - c) o_in + B (c real, B complex)
//   2. k_dcb_scan    one CTA composes the segment maps into the offset at every segment start (from the carried state)
//   3. k_dcb_apply   every thread re-derives the offset at the start of its DCB_RUN-sample run from the thread / warp /
//                    block composition of the same maps, then walks its run with the reference's own two statements
//                    (so inside a run the arithmetic is the reference's); the conjugate is applied on the store.
// The last offset goes back to the state word for the next chunk.
#pragma once

#define DCB_RUN 16                       // consecutive samples one thread walks
#define DCB_THREADS 256
#define DCB_SEG (DCB_RUN * DCB_THREADS)  // samples per CTA


// affine map over complex numbers with a real slope, o -> (1 - c) o + b.  The slope is kept as its COMPLEMENT c: with
// rate = 50 / fs ~ 5e-7 the slope of one step is 1 - 5e-7, which fp32 can only hold to 6 % of its distance from one;
// c = 5e-7 (and c1 + c2 - c1 c2 for a composition) keeps the blocker's time constant to full precision.
struct DcbMap { float c; float2 b; };
__device__ __forceinline__ DcbMap dcb_identity() {
    DcbMap m;
    m.c = 0.0f;
    m.b = make_float2(0.0f, 0.0f);
    return m;
}
__device__ __forceinline__ float2 dcb_apply_map(const DcbMap& m, float2 o) {
    return make_float2(o.x - m.c * o.x + m.b.x, o.y - m.c * o.y + m.b.y);
}
__device__ __forceinline__ DcbMap dcb_compose(const DcbMap& first, const DcbMap& then) {      // then(first(o))
    DcbMap m;
    m.c = first.c + then.c - first.c * then.c;
    m.b = dcb_apply_map(then, first.b);
    return m;
}
// one blocker step with input x appended to a map: o'' = o' + (x - o') r
__device__ __forceinline__ void dcb_step(DcbMap& m, float2 x, float r) {
    m.c = m.c + r - m.c * r;
    m.b = make_float2(m.b.x + (x.x - m.b.x) * r, m.b.y + (x.y - m.b.y) * r);
}
__device__ __forceinline__ float2 dcb_load(const DcbParams& p, long long i) {
    if (p.fmt == FMT_CF32) { return load_iq<FMT_CF32>(p.in, i, 0.0f); }
    if (p.fmt == FMT_CS16) { return load_iq<FMT_CS16>(p.in, i, p.in_scale); }
    return load_iq<FMT_CS8>(p.in, i, p.in_scale);
}
// map of the run [i0, i1): walking o through it with the blocker's update
__device__ __forceinline__ DcbMap dcb_run_map(const DcbParams& p, long long i0, long long i1) {
    DcbMap m = dcb_identity();
    for (long long i = i0; i < i1; i++) { dcb_step(m, dcb_load(p, i), p.rate); }
    return m;
}
// inclusive scan of per-thread maps over the CTA (thread order); returns the EXCLUSIVE prefix of this thread and, through
// `total`, the composition of the whole CTA
__device__ __forceinline__ DcbMap dcb_block_exclusive(DcbMap mine, DcbMap* warp_tot, DcbMap& total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    DcbMap inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        DcbMap o;
        o.c = __shfl_up_sync(0xffffffffu, inc.c, d);
        o.b.x = __shfl_up_sync(0xffffffffu, inc.b.x, d);
        o.b.y = __shfl_up_sync(0xffffffffu, inc.b.y, d);
        if (lane >= d) { inc = dcb_compose(o, inc); }
    }
    if (lane == 31) { warp_tot[w] = inc; }
    __syncthreads();
    DcbMap pre = dcb_identity();                  // composition of the warps before this one
    DcbMap all = pre;
    for (int k = 0; k < DCB_THREADS / 32; k++) {
        if (k == w) { pre = all; }
        all = dcb_compose(all, warp_tot[k]);
    }
    total = all;
    // exclusive prefix inside the warp = inclusive of the previous lane
    DcbMap ex;
    ex.c = __shfl_up_sync(0xffffffffu, inc.c, 1);
    ex.b.x = __shfl_up_sync(0xffffffffu, inc.b.x, 1);
    ex.b.y = __shfl_up_sync(0xffffffffu, inc.b.y, 1);
    if (lane == 0) { ex = dcb_identity(); }
    return dcb_compose(pre, ex);
}

__global__ void __launch_bounds__(DCB_THREADS) k_dcb_reduce(const __grid_constant__ DcbParams p) {
    __shared__ DcbMap wt[DCB_THREADS / 32];
    const long long s0 = (long long)blockIdx.x * DCB_SEG + (long long)threadIdx.x * DCB_RUN;
    const long long lo = s0 < p.count ? s0 : (long long)p.count;
    const long long hi = s0 + DCB_RUN < p.count ? s0 + DCB_RUN : (long long)p.count;
    DcbMap mine = dcb_run_map(p, lo, hi > lo ? hi : lo);          // threads past the end of the chunk: the identity
    DcbMap total;
    dcb_block_exclusive(mine, wt, total);
    if (threadIdx.x == 0) { p.segA[blockIdx.x] = total.c; p.segB[blockIdx.x] = total.b; }
}

// one CTA: segB[s] <- offset at the start of segment s; state <- offset after the last segment
__global__ void __launch_bounds__(1024) k_dcb_scan(const __grid_constant__ DcbParams p) {
    __shared__ DcbMap part[1024];
    const int t = threadIdx.x;
    const int per = (p.nseg + 1023) / 1024;
    const int lo = t * per, hi = min(p.nseg, lo + per);
    DcbMap m = dcb_identity();
    for (int s = lo; s < hi; s++) {
        DcbMap g;
        g.c = p.segA[s]; g.b = p.segB[s];
        m = dcb_compose(m, g);
    }
    part[t] = m;
    __syncthreads();
    // Hillis-Steele over the 1024 partial maps
    for (int d = 1; d < 1024; d <<= 1) {
        DcbMap o = part[t];
        if (t >= d) { o = dcb_compose(part[t - d], o); }
        __syncthreads();
        part[t] = o;
        __syncthreads();
    }
    float2 o = p.state[0];
    if (t > 0) { o = dcb_apply_map(part[t - 1], o); }
    for (int s = lo; s < hi; s++) {
        DcbMap g;
        g.c = p.segA[s]; g.b = p.segB[s];
        p.segB[s] = o;
        o = dcb_apply_map(g, o);
    }
    __syncthreads();
    if (hi == p.nseg && lo < hi) { p.state[0] = o; }
}

__global__ void __launch_bounds__(DCB_THREADS) k_dcb_apply(const __grid_constant__ DcbParams p) {
    __shared__ DcbMap wt[DCB_THREADS / 32];
    const long long s0 = (long long)blockIdx.x * DCB_SEG + (long long)threadIdx.x * DCB_RUN;
    long long s1 = s0 + DCB_RUN;
    if (s1 > p.count) { s1 = p.count; }
    float2 x[DCB_RUN];
#pragma unroll
    for (int k = 0; k < DCB_RUN; k++) { x[k] = (s0 + k < s1) ? dcb_load(p, s0 + k) : make_float2(0.0f, 0.0f); }
    float2 o = make_float2(0.0f, 0.0f);
    if (p.dc_on) {
        DcbMap mine = dcb_identity();
#pragma unroll
        for (int k = 0; k < DCB_RUN; k++) {
            if (s0 + k < s1) { dcb_step(mine, x[k], p.rate); }
        }
        DcbMap total;
        const DcbMap ex = dcb_block_exclusive(mine, wt, total);
        o = dcb_apply_map(ex, p.segB[blockIdx.x]);                // segB: offset at the segment start (k_dcb_scan)
    }
    const float sgn = p.conj_on ? -1.0f : 1.0f;
#pragma unroll
    for (int k = 0; k < DCB_RUN; k++) {
        if (s0 + k < s1) {
            float2 y = x[k];
            if (p.dc_on) {
                // the reference's two statements (dc_blocker.h:56-57), explicit roundings
                y = make_float2(__fsub_rn(x[k].x, o.x), __fsub_rn(x[k].y, o.y));
                o = make_float2(__fadd_rn(o.x, __fmul_rn(y.x, p.rate)), __fadd_rn(o.y, __fmul_rn(y.y, p.rate)));
            }
            p.out[s0 + k] = make_float2(y.x, sgn * y.y);
        }
    }
}

cudaError_t launch_preproc(const DcbParams& p, cudaStream_t s, int* nlaunch) {
    if (p.count <= 0) { return cudaSuccess; }
    const int nseg = (p.count + DCB_SEG - 1) / DCB_SEG;
    if (p.dc_on) {
        k_dcb_reduce<<<nseg, DCB_THREADS, 0, s>>>(p);
        k_dcb_scan<<<1, 1024, 0, s>>>(p);
        if (nlaunch) { *nlaunch += 2; }
    }
    k_dcb_apply<<<nseg, DCB_THREADS, 0, s>>>(p);
    if (nlaunch) { *nlaunch += 1; }
    return cudaGetLastError();
}
