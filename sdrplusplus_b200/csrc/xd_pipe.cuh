// sdrplusplus_b200/csrc/xd_pipe.cuh -- stage 1 (frequency translate + first decimating FIR of every VFO),
// persistent double-buffered variant ("s1" = 3, the default).  Included by kernels.cu.
//
// One CTA per SM walks over tiles of MT decimated output positions.  While the warps convolve tile t out of
// shared-memory buffer b, the raw IQ of tile t+1 streams into buffer b^1 with cp.async (8-byte LDGSTS whose
// shared-memory destination de-interleaves the samples by decimation phase: X[r][j] = x((J0+j)*D + r + org)),
// so HBM latency is hidden behind the FMAs instead of stalling every warp of the CTA (the single-buffered
// k_xd_tile spent most of its time in long-scoreboard stalls, profiles/r01_xd_tile_v1.txt).
//
// Work split per tile: (strip of 128 outputs) x (group of 4 VFOs) x (RS-way split of the D phases); every warp
// owns one such task: 4 outputs x 4 VFOs per lane in registers, taps read as broadcast LDS.128 (two VFOs per
// load), inner product as packed f32x2 FMAs with the tap as scalar-broadcast operand.  Partial sums of an RS
// split are exchanged through shared memory.
#pragma once

#define XP_VR 4
#define XP_RM 4

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

struct XpGeom {
    int MT, JP, QPC, ntiles, org, RS, logD;
    int p_alias;      // 1: the RS exchange buffer lives in the tile buffer just consumed (single task round only)
    int single;       // 1: one tile buffer per CTA (several CTAs per SM overlap each other instead)
    long long jmin;
};

template <int FMT, int QC, int NT>
__global__ void __launch_bounds__(NT, 1) k_xd_pipe(const __grid_constant__ XdParams p, const XpGeom g) {
    extern __shared__ __align__(16) float2 smem[];
    const int D = p.D, JP = g.JP, MT = g.MT, QPC = g.QPC;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const int gl = QPC * D;
    const int ngroups = (p.nslots + XP_VR - 1) / XP_VR;
    float2* Xb0 = smem;
    float2* Xb1 = g.single ? smem : smem + (size_t)D * JP;
    float2* G = smem + (size_t)(g.single ? 1 : 2) * D * JP;   // [ngroups][gl][XP_VR]
    float2* TB = G + (size_t)ngroups * gl * XP_VR;            // [njobs][MT] phase ramp e^{j W_v D k} + [njobs] per-tile base
    float2* BASE = TB + (size_t)p.njobs * MT;
    int* CJ = reinterpret_cast<int*>(BASE + B200_BATCH);      // [B200_BATCH] block offset c of every job
    int* JN = CJ + B200_BATCH;                                 // [B200_BATCH] n_out of every job
    float2** JOUT = reinterpret_cast<float2**>(BASE + 2 * B200_BATCH);   // [B200_BATCH] output pointers
    float2* P = BASE + 3 * B200_BATCH;                         // [nwarps][32][32] partial sums A,B (no alias)

    // ---- taps: G[grp][k][vv] = gpad_v[(D-1-s_v) + k], zero for VFO slots beyond njobs ----
    for (int idx = tid; idx < ngroups * gl * XP_VR; idx += nthr) {
        int vv = idx % XP_VR, k = (idx / XP_VR) % gl, grp = idx / (XP_VR * gl);
        int slot = grp * XP_VR + vv;
        float2 t = make_float2(0.0f, 0.0f);
        if (slot < p.nslots) {
            const XdJob& Jv = p.job[p.slot_a[slot]];
            int a = Jv.offset - (Jv.T - 1) - g.org;
            int s = ((a % D) + D) % D;
            t = __ldg(Jv.gpad + (D - 1 - s) + k);
        }
        G[idx] = t;
    }

    if (tid < p.njobs) {
        const XdJob& Jv = p.job[tid];
        const int a = Jv.offset - (Jv.T - 1) - g.org;
        const int sft = ((a % D) + D) % D;
        CJ[tid] = (a - sft) / D;
        JN[tid] = Jv.n_out;
        JOUT[tid] = Jv.out;
    }
    // phase ramp of every job over one tile: exact u64 phase, one sincospi per entry, once per CTA
    for (int idx = tid; idx < p.njobs * MT; idx += nthr) {
        const int v = idx / MT, k = idx - v * MT;
        TB[idx] = phasor_u64(p.job[v].w * (unsigned long long)((long long)k * D));
    }
    const int ntile_samples = D * (MT + QPC);
    const int dmask = D - 1;
    auto issue = [&](int tile, float2* X) {
        const long long ibase = (g.jmin + (long long)tile * MT) * D + g.org;
        if (FMT == FMT_CF32 && ibase >= 0 && ibase + ntile_samples <= (long long)p.count && (nthr & dmask) == 0) {
            // interior tile: every sample comes straight from the chunk; the thread's decimation phase r is fixed,
            // only the block index j advances -> one add per element
            const int r = tid & dmask;
            const int jstep = nthr >> g.logD;
            float2* dst = X + r * JP + (tid >> g.logD);
            const float2* src = reinterpret_cast<const float2*>(p.in) + ibase + tid;
            for (int idx = tid; idx < ntile_samples; idx += nthr) {
                cp_async8(dst, src);
                dst += jstep;
                src += nthr;
            }
        }
        else if (FMT != FMT_CF32 && ibase >= 0 && ibase + ntile_samples <= (long long)p.count && (nthr & dmask) == 0) {
            // interior tile of an integer stream: same walk, converted in registers (8 loads in flight per thread)
            const int r = tid & dmask;
            const int jstep = nthr >> g.logD;
            float2* dst = X + r * JP + (tid >> g.logD);
            long long i = ibase + tid;
#pragma unroll 8
            for (int idx = tid; idx < ntile_samples; idx += nthr) {
                *dst = load_iq<FMT>(p.in, i, p.in_scale);
                dst += jstep;
                i += nthr;
            }
        }
        else {
            for (int idx = tid; idx < ntile_samples; idx += nthr) {
                const int j = idx >> g.logD, r = idx & dmask;
                const long long i = ibase + idx;
                float2* dst = X + r * JP + j;
                if (FMT == FMT_CF32 && i >= 0 && i < p.count) {
                    cp_async8(dst, reinterpret_cast<const float2*>(p.in) + i);
                }
                else {
                    *dst = load_x<FMT>(p, i);
                }
            }
        }
        cp_async_commit();
    };

    const int nstrips = MT / (32 * XP_RM);
    const int RS = g.RS;
    const int ntasks = nstrips * ngroups * RS;
    const int rper = D / RS;
    constexpr int NP = XP_RM / 2;
    constexpr int WN = (QC + 2) & ~1;             // window samples per pair: QC+1 needed, loaded as LDS.128 pairs

    int tile = blockIdx.x, buf = 0;
    if (tile < g.ntiles && !g.single) { issue(tile, Xb0); }
    for (; tile < g.ntiles; tile += gridDim.x, buf ^= 1) {
        const int next = tile + gridDim.x;
        float2* X = buf ? Xb1 : Xb0;
        if (g.single) {
            issue(tile, X);
            cp_async_wait<0>();
        }
        else if (next < g.ntiles) {
            issue(next, buf ? Xb0 : Xb1);
            cp_async_wait<1>();
        }
        else {
            cp_async_wait<0>();
        }
        const long long J0 = g.jmin + (long long)tile * MT;
        if (tid < p.njobs) {
            // phase of this job at the tile's first output position (m = J0 - c): base * ramp[k] gives output J0 + k
            const XdJob& Jv = p.job[tid];
            const int a0 = Jv.offset - (Jv.T - 1);
            const int a = a0 - g.org;
            const int s = ((a % D) + D) % D;
            const long long c = (a - s) / D;
            const long long im0 = (long long)a0 + (J0 - c) * D;
            BASE[tid] = phasor_u64(Jv.phase0 + Jv.w * (unsigned long long)im0);
        }
        __syncthreads();

        for (int task0 = 0; task0 < ntasks; task0 += nwarps) {
            const int task = task0 + warp;
            const bool active = task < ntasks;
            const int half = task % RS, sg = task / RS;
            const int strip = sg % nstrips, grp = sg / nstrips;
            float2 A[NP][2][XP_VR], B[NP][2][XP_VR];
            const int jl0 = strip * 32 * XP_RM + 2 * lane;
            if (active) {
#pragma unroll
                for (int pi = 0; pi < NP; pi++)
#pragma unroll
                    for (int o = 0; o < 2; o++)
#pragma unroll
                        for (int v = 0; v < XP_VR; v++) { A[pi][o][v] = make_float2(0.f, 0.f); B[pi][o][v] = make_float2(0.f, 0.f); }
                const float2* Gg = G + (size_t)grp * gl * XP_VR;
                for (int r = half * rper; r < (half + 1) * rper; r++) {
                    const float2* row = X + r * JP;
                    for (int qc = 0; qc < QPC; qc += QC) {
                        float2 xs[NP][WN];
#pragma unroll
                        for (int pi = 0; pi < NP; pi++) {
                            const float4* src = reinterpret_cast<const float4*>(row + jl0 + 64 * pi + qc);
#pragma unroll
                            for (int u = 0; u < WN / 2; u++) {
                                float4 t = src[u];
                                xs[pi][2 * u] = make_float2(t.x, t.y);
                                xs[pi][2 * u + 1] = make_float2(t.z, t.w);
                            }
                        }
#pragma unroll
                        for (int q = 0; q < QC; q++) {
                            const float4* tp = reinterpret_cast<const float4*>(Gg + (size_t)((qc + q) * D + r) * XP_VR);
#pragma unroll
                            for (int vp = 0; vp < XP_VR / 2; vp++) {
                                const float4 t2 = tp[vp];            // taps of two slots, broadcast
                                const float2 gA = make_float2(t2.x, t2.y), gB = make_float2(t2.z, t2.w);
#pragma unroll
                                for (int pi = 0; pi < NP; pi++)
#pragma unroll
                                    for (int o = 0; o < 2; o++) {
                                        A[pi][o][2 * vp] = ffma2(make_float2(gA.x, gA.x), xs[pi][q + o], A[pi][o][2 * vp]);
                                        B[pi][o][2 * vp] = ffma2(make_float2(gA.y, gA.y), xs[pi][q + o], B[pi][o][2 * vp]);
                                        A[pi][o][2 * vp + 1] = ffma2(make_float2(gB.x, gB.x), xs[pi][q + o], A[pi][o][2 * vp + 1]);
                                        B[pi][o][2 * vp + 1] = ffma2(make_float2(gB.y, gB.y), xs[pi][q + o], B[pi][o][2 * vp + 1]);
                                    }
                            }
                        }
                    }
                }
            }
            float2* Pb = g.p_alias ? X : P;
            {
                // Every warp publishes its 16 (A,B) pairs; the 16 (output, slot) combinations of a task are then dealt
                // round-robin to its RS warps, which sum the RS partials and do the epilogue (phase rotation + store).
                // Going through shared memory keeps this loop small (runtime indices) and spreads it over all warps.
                if (g.p_alias) { __syncthreads(); }         // every warp is done reading X before it is reused
                if (active) {
                    float2* dst = Pb + (size_t)warp * 32 * 32 + lane;
#pragma unroll
                    for (int pi = 0; pi < NP; pi++)
#pragma unroll
                        for (int o = 0; o < 2; o++)
#pragma unroll
                            for (int v = 0; v < XP_VR; v++) {
                                dst[(((pi * 2 + o) * XP_VR + v) * 2 + 0) * 32] = A[pi][o][v];
                                dst[(((pi * 2 + o) * XP_VR + v) * 2 + 1) * 32] = B[pi][o][v];
                            }
                }
                __syncthreads();
            }
            if (active) {
                const float2* grpP = Pb + (size_t)(warp - half) * 32 * 32 + lane;      // first warp of this task's group
                for (int ci = half; ci < NP * 2 * XP_VR; ci += RS) {
                    const int v = ci & (XP_VR - 1), o = (ci / XP_VR) & 1, pi = ci / (2 * XP_VR);
                    const int slot = grp * XP_VR + v;
                    if (slot >= p.nslots) { continue; }
                    float2 Av = make_float2(0.f, 0.f), Bv = make_float2(0.f, 0.f);
                    for (int h = 0; h < RS; h++) {
                        const float2 ta = grpP[(size_t)h * 32 * 32 + (ci * 2 + 0) * 32];
                        const float2 tb = grpP[(size_t)h * 32 * 32 + (ci * 2 + 1) * 32];
                        Av.x += ta.x; Av.y += ta.y;
                        Bv.x += tb.x; Bv.y += tb.y;
                    }
                    const int jl = jl0 + 64 * pi + o;
                    const int ja = p.slot_a[slot], jb = p.slot_b[slot];
                    {
                        const long long m = J0 + jl - (long long)CJ[ja];
                        if (m >= 0 && m < JN[ja]) {
                            const float2 ph = cmulf(BASE[ja], TB[(size_t)ja * MT + jl]);
                            JOUT[ja][m] = cmulf(make_float2(Av.x - Bv.y, Av.y + Bv.x), ph);
                        }
                    }
                    if (jb >= 0) {
                        const long long m = J0 + jl - (long long)CJ[jb];
                        if (m >= 0 && m < JN[jb]) {
                            const float2 ph = cmulf(BASE[jb], TB[(size_t)jb * MT + jl]);
                            JOUT[jb][m] = cmulf(make_float2(Av.x + Bv.y, Av.y - Bv.x), ph);     // conjugate taps
                        }
                    }
                }
            }
            if (task0 + nwarps < ntasks) { __syncthreads(); }    // P is reused by the next round
        }
        __syncthreads();       // everyone is done with X before the next iteration's prefetch overwrites it
    }
}
