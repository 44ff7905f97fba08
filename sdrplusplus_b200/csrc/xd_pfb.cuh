// sdrplusplus_b200/csrc/xd_pfb.cuh -- stage 1 for VFO plans whose offsets are commensurate with the sample rate
// ("s1" = 7, the default when it applies; falls back to k_xd_pipe).  Included by kernels.cu after xd_pipe.cuh.
//
// Stage 1 computes, for every VFO v,   y_v[m] = e^{j phi_v(i_m)} * sum_k h[k] e^{j w_v k} x[i_m + k]
// (FrequencyXlator folded into the first DecimatingFIR: frequency_xlator.h:43-50, decimating_fir.h:45-68).
// When e^{j w_v PS} = sigma (= +1 or -1) for EVERY VFO of the launch -- offsets on a grid of fs / (2 PS), e.g. the
// +-5/15/25/35 MHz plan of BASELINE config 2 on a 100 MS/s stream: PS = 10, sigma = -1 -- the tap phasor only
// depends on k mod PS:   e^{j w_v k} = sigma^floor(k / PS) * e^{j w_v (k mod PS)},   so
//     S_a[m]  = sum_{k = a (mod PS)} sigma^floor(k/PS) h[k] x[i_m + k]        (a < PS; REAL taps, shared by all VFOs)
//     y_v[m]  = e^{j phi_v(i_m)} * sum_a e^{j w_v a} S_a[m]                     (PS complex MACs per VFO)
// i.e. a polyphase filter bank: 2 real FMAs per tap and sample for ALL VFOs together instead of 4 per VFO, which
// moves stage 1 from the fp32 roof to the HBM roof.  The identity is exact for offsets that are exact multiples of
// fs / (2 PS); the reference's phase increment is the angle of an fp32-rounded phasor, a few 1e-8 rad off that grid:
// the host accepts a plan when the phase drift across one tap window stays below 1e-6 rad (Scheduler::run), and the
// output phase e^{j phi_v} keeps using the exact 64-bit phase, so nothing accumulates.
//
// Tile = 256 outputs, 64 per warp: a lane walks all D decimation phases of the de-interleaved tile for its 2 outputs
// (PS complex accumulators each), so no partial sums cross warps; VFOs at +f / -f share the combination sums
// (XdParams slots).  The per-output phase e^{j phi_v} comes from the exact 64-bit phase: a per-tile table of 16 coarse
// phases per VFO times a 16-entry fine ramp.
#pragma once

#define PFB_MT 256
template <int LOGD, int QC, int PS>
__global__ void __launch_bounds__(128, 3) k_xd_pfb(const __grid_constant__ XdParams p, const XpGeom g, const int fmt) {
    extern __shared__ __align__(16) float2 smem[];
    constexpr int D = 1 << LOGD, MT = PFB_MT;
    constexpr int WN = (QC + 2) & ~1;             // window samples per output pair (QC + 1 needed, loaded as LDS.128)
    constexpr int GQ = (QC + 3) & ~3;             // taps per phase row, padded to whole LDS.128
    const int JP = g.JP;
    const int tid = threadIdx.x, lane = tid & 31, h = tid >> 5;
    float2* X = smem;                                               // [D][JP]
    float* Gs = reinterpret_cast<float*>(smem + (size_t)D * JP);    // [D][GQ] signed real taps
    float2* C = reinterpret_cast<float2*>(Gs + D * GQ);             // [njobs][PS]  e^{j w_v a}
    float2* TL = C + (size_t)B200_BATCH * PS;                        // [njobs][16]    e^{j w_v D j}
    float2* PHT = TL + (size_t)B200_BATCH * 16;                      // [njobs][MT/16] per tile: base phase * coarse ramp
    float2* BASE = PHT + (size_t)B200_BATCH * (MT / 16);
    int* CJ = reinterpret_cast<int*>(BASE + B200_BATCH);
    int* JN = CJ + B200_BATCH;
    float2** JOUT = reinterpret_cast<float2**>(BASE + 2 * B200_BATCH);

    // ---- tables (once per CTA) ----
    {
        const XdJob& J0j = p.job[0];
        for (int idx = tid; idx < D * GQ; idx += 128) {
            const int r = idx / GQ, q = idx - r * GQ;
            const int k = q * D + r;
            float t = (q < QC && k < J0j.T) ? __ldg(J0j.h + k) : 0.0f;
            if (p.pfb_sigma < 0 && ((k / PS) & 1)) { t = -t; }
            Gs[idx] = t;
        }
        for (int idx = tid; idx < p.njobs * PS; idx += 128) {
            const int v = idx / PS, a = idx - v * PS;
            C[idx] = phasor_u64(p.job[v].w * (unsigned long long)a);
        }
        if (tid < p.njobs) {
            const XdJob& Jv = p.job[tid];
            const int a = Jv.offset - (Jv.T - 1) - g.org;            // == 0 (mod D): every job shares job 0's alignment
            CJ[tid] = a / D;
            JN[tid] = Jv.n_out;
            JOUT[tid] = Jv.out;
        }
        for (int idx = tid; idx < p.njobs * 16; idx += 128) {
            const int v = idx >> 4, j = idx & 15;
            TL[idx] = phasor_u64(p.job[v].w * (unsigned long long)((long long)j * D));
        }
    }
    __syncthreads();                                // CJ is read by other threads when the first tile's phase table is built
    const int ntile_samples = D * (MT + QC);
    const int jlw = h * 64 + 2 * lane;              // this lane's first output of the tile

    for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
        const long long J0 = g.jmin + (long long)tile * MT;
        // ---- tile fill: X[r][j] = x((J0 + j) * D + org + r) ----
        {
            const long long ibase = J0 * D + g.org;
            const bool interior = ibase >= 0 && ibase + ntile_samples <= (long long)p.count;
            const int r = tid & (D - 1);
            constexpr int jstep = 128 >> LOGD;
            if (interior && D <= 128) {
                float2* dst = X + r * JP + (tid >> LOGD);
                if (fmt == FMT_CF32) {
                    const float2* src = reinterpret_cast<const float2*>(p.in) + ibase + tid;
#pragma unroll 4
                    for (int idx = tid; idx < ntile_samples; idx += 128) { cp_async8(dst, src); dst += jstep; src += 128; }
                }
                else if (fmt == FMT_CS16) {
                    long long i = ibase + tid;
#pragma unroll 8
                    for (int idx = tid; idx < ntile_samples; idx += 128) { *dst = load_iq<FMT_CS16>(p.in, i, p.in_scale); dst += jstep; i += 128; }
                }
                else {
                    long long i = ibase + tid;
#pragma unroll 8
                    for (int idx = tid; idx < ntile_samples; idx += 128) { *dst = load_iq<FMT_CS8>(p.in, i, p.in_scale); dst += jstep; i += 128; }
                }
            }
            else {
                for (int idx = tid; idx < ntile_samples; idx += 128) {
                    const long long i = ibase + idx;
                    float2 v;
                    if (fmt == FMT_CF32) { v = load_x<FMT_CF32>(p, i); }
                    else if (fmt == FMT_CS16) { v = load_x<FMT_CS16>(p, i); }
                    else { v = load_x<FMT_CS8>(p, i); }
                    X[(idx & (D - 1)) * JP + (idx >> LOGD)] = v;
                }
            }
            cp_async_commit();
            cp_async_wait<0>();
        }
        {
            // per-tile phase of every job at the tile's first output, times the coarse ramp: 16 entries per job
            const int v = tid >> 4, i = tid & 15;                    // MT / 16 == 16 entries, 8 jobs per pass of 128 threads
            for (int vv = v; vv < p.njobs; vv += 8) {
                const XdJob& Jv = p.job[vv];
                const int a0 = Jv.offset - (Jv.T - 1);
                const long long im0 = (long long)a0 + (J0 - (long long)CJ[vv]) * D;
                // centre the (tiny) drift of e^{j w PS n} against sigma^n on the middle of the tap window
                unsigned long long dr = Jv.w * (unsigned long long)PS;
                if (p.pfb_sigma < 0) { dr -= 0x8000000000000000ULL; }
                const long long corr = (long long)dr * (long long)(Jv.T / (2 * PS));
                PHT[vv * (MT / 16) + i] = phasor_u64(Jv.phase0 + Jv.w * (unsigned long long)(im0 + (long long)i * 16 * D) + (unsigned long long)corr);
            }
        }
        __syncthreads();

        // ---- accumulate: all D phases, 2 outputs x PS accumulators per lane ----
        float2 S[2][PS];
#pragma unroll
        for (int o = 0; o < 2; o++)
#pragma unroll
            for (int a = 0; a < PS; a++) { S[o][a] = make_float2(0.f, 0.f); }
#pragma unroll
        for (int r = 0; r < D; r++) {
            // JP is odd (conflict-free de-interleaving stores): odd rows start 8 bytes off a 16-byte boundary, their
            // window is loaded from one column earlier (sh = 1)
            constexpr int WNO = (QC + 3) & ~1;
            const int sh = r & 1;
            const float4* src = reinterpret_cast<const float4*>(X + r * JP + jlw - sh);
            float2 xs[WNO];
#pragma unroll
            for (int u = 0; u < WNO / 2; u++) {
                if (2 * u < WN + 2 * sh) {
                    const float4 t = src[u];
                    xs[2 * u] = make_float2(t.x, t.y);
                    xs[2 * u + 1] = make_float2(t.z, t.w);
                }
            }
            float gq[GQ];
            const float4* gp = reinterpret_cast<const float4*>(Gs + r * GQ);
#pragma unroll
            for (int u = 0; u < GQ / 4; u++) {
                const float4 t = gp[u];
                gq[4 * u] = t.x; gq[4 * u + 1] = t.y; gq[4 * u + 2] = t.z; gq[4 * u + 3] = t.w;
            }
#pragma unroll
            for (int q = 0; q < QC; q++) {
                const int a = (q * D + r) % PS;              // compile-time after unrolling
#pragma unroll
                for (int o = 0; o < 2; o++) { S[o][a] = ffma2(make_float2(gq[q], gq[q]), xs[q + o + sh], S[o][a]); }
            }
        }

        // ---- combine per slot (a VFO, or a +f / -f pair sharing A = sum cos*S and B = sum sin*S), rotate, store ----
        // every job of a filter-bank launch has the same alignment and length: one bounds test per output
        const long long m0 = J0 + jlw - (long long)CJ[0];
        const bool ok0 = m0 >= 0 && m0 < JN[0], ok1 = (m0 + 1) >= 0 && (m0 + 1) < JN[0];
        const int ih = jlw >> 4, il = jlw & 15;                 // jlw is even: both outputs share the coarse ramp entry
        for (int sl = 0; sl < p.nslots; sl++) {
            const int ja = p.slot_a[sl], jb = p.slot_b[sl];
            float2 A0 = make_float2(0.f, 0.f), B0 = A0, A1 = A0, B1 = A0;
            const float2* cv = C + ja * PS;
#pragma unroll
            for (int a = 0; a < PS; a++) {
                const float2 c = cv[a];
                A0 = ffma2(make_float2(c.x, c.x), S[0][a], A0);
                B0 = ffma2(make_float2(c.y, c.y), S[0][a], B0);
                A1 = ffma2(make_float2(c.x, c.x), S[1][a], A1);
                B1 = ffma2(make_float2(c.y, c.y), S[1][a], B1);
            }
            {
                const float2 pc = PHT[ja * (MT / 16) + ih];
                const float2 p0 = cmulf(pc, TL[ja * 16 + il]), p1 = cmulf(pc, TL[ja * 16 + il + 1]);
                float2* out = JOUT[ja] + m0;
                if (ok0) { out[0] = cmulf(make_float2(A0.x - B0.y, A0.y + B0.x), p0); }
                if (ok1) { out[1] = cmulf(make_float2(A1.x - B1.y, A1.y + B1.x), p1); }
            }
            if (jb >= 0) {
                const float2 pc = PHT[jb * (MT / 16) + ih];
                const float2 p0 = cmulf(pc, TL[jb * 16 + il]), p1 = cmulf(pc, TL[jb * 16 + il + 1]);
                float2* out = JOUT[jb] + m0;
                if (ok0) { out[0] = cmulf(make_float2(A0.x + B0.y, A0.y - B0.x), p0); }      // conjugate coefficients
                if (ok1) { out[1] = cmulf(make_float2(A1.x + B1.y, A1.y - B1.x), p1); }
            }
        }
        __syncthreads();                                       // tile consumed before the next fill
    }
}

