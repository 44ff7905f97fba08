// sdrplusplus_b200/csrc/xd_tma.cuh -- stage 1 in filter-bank form (see xd_pfb.cuh for the identity) with the raw IQ
// tile brought in by the TMA engine ("s1" = 8, the default for cf32 chunks when the plan allows it; falls back to
// k_xd_pfb / k_xd_pipe).  Included by kernels.cu after xd_pfb.cuh.
//
// Replaces FrequencyXlator::process + the first DecimatingFIR of every VFO (frequency_xlator.h:43-50,
// decimating_fir.h:45-68) and the Splitter fan-out in front of them (splitter.h:46-61): the chunk is read from HBM once.
//
// Why a second filter-bank kernel: k_xd_pfb de-interleaves its tile by decimation phase with 8-byte cp.async copies,
// fills and computes in turn, and was latency-bound at a third of the HBM roof.  Here
//   * the chunk is described to the TMA engine as a 2-D tensor of "super-rows" (2*D samples = two decimation blocks);
//     one tile = NSEG*2 boxes of {16 samples, NR super-rows} with the 128-byte swizzle, issued by ONE thread, landing in
//     shared memory as [segment][block parity][row][128 B].  No thread touches the data on its way in.
//   * a persistent CTA per SM runs a ring of XT_STAGES tiles guarded by full / empty mbarriers: a producer warp keeps
//     two tiles in flight while MT/64 consumer warps filter the third, so HBM latency never reaches the FMAs.
//   * a lane owns two ADJACENT outputs; their windows (QC+1 blocks) sit in consecutive rows of the two parity regions,
//     so every 16-byte window load of a warp is conflict-free through the hardware swizzle and feeds four packed FMAs;
//     the real taps come from the kernel parameter block (constant bank), not from shared memory.
// Rows are cut at a 128-byte aligned origin org' <= org; the s = org - org' samples in front of a window are covered
// by s leading zero taps: the accumulator class of window position p is p mod PS, the tap it carries is k = p - s.
#pragma once
#include <cuda.h>

#define XT_STAGES NST        // ring depth, a template parameter of the kernel (2 or 3)
#define XT_MAXTAPS 512

struct XtGeom {
    float g[XT_MAXTAPS];      // signed real taps by window position p = q*D + r:  sigma^floor(k/PS) h[k], k = p - s
    long long jmin;           // block index (relative to the row origin) of the first tile's first output, even
    int ntiles;
    int org;                  // sample index of row 0, phase 0 (chunk-relative; may be negative: history)
    int s;                    // leading zero taps
    int cj;                   // window of output m starts in block cj + m
    int rows_tma;             // super-rows the tensor map covers (all inside the chunk)
    int diag;                 // measurement only (option "s1_diag"): 1 = tiles are loaded but not filtered (what the TMA ring alone
                              // sustains), 2 = tiles are filtered but never loaded (what the consumer warps alone sustain); results are garbage
};

__device__ __forceinline__ void xt_mbar_init(unsigned bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void xt_mbar_wait(unsigned bar, unsigned parity) {
    asm volatile("{\n.reg .pred p;\nXT_WAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra XT_WAIT_%=;\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void xt_mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void xt_mbar_expect(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
// The raw chunk is read once and never again by this kernel: its lines are marked evict-first in L2, so that streaming 128 MiB
// through does not push out what the kernels of the other streams come back for (stage outputs, FFT work buffers).
__device__ __forceinline__ unsigned long long xt_policy_evict_first() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void xt_tma_2d(unsigned dst, const CUtensorMap* tm, unsigned bar, int c0, int c1, unsigned long long pol) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;\n"
                 ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(tm)), "r"(bar), "r"(c0), "r"(c1), "l"(pol) : "memory");
}

template <int LOGD, int QC, int MT>
struct XtLay {
    static constexpr int D = 1 << LOGD;
    static constexpr int NSEG = D / 16;                                  // 128-byte segments per decimation block
    static constexpr int NEED = (MT + QC + 2) / 2;                       // super-rows a tile reads
    static constexpr int NR = (NEED + 7) & ~7;                           // rows per region (whole swizzle atoms)
    static constexpr int REGION = NR * 128;
    static constexpr int STAGE = NSEG * 2 * REGION;
    static constexpr int NW = MT / 64;                                   // consumer warps
};

// decimation-phase pairs [RP0, RP1) of a tile -- they all lie in ONE 128-byte segment (8 pairs per segment), whose two parity
// regions start at sb -- into the 2 x PS class sums of a lane.  Every index is a compile-time constant: taps from the constant
// bank, accumulators in registers.
template <int LOGD, int QC, int PS, int REGION, int RP0, int RP1, int NQH>
__device__ __forceinline__ void xt_accumulate(float2 (&S)[2][PS], const XtGeom& g, unsigned sb, const unsigned (&aoff)[NQH]) {
    constexpr int D = 1 << LOGD;
    static_assert((RP0 >> 3) == ((RP1 - 1) >> 3), "one segment per call");
#pragma unroll
    for (int rp = RP0; rp < RP1; rp++) {
#pragma unroll
        for (int q = 0; q <= QC; q++) {
            const unsigned addr = sb + (unsigned)((q & 1) * REGION) + (aoff[q >> 1] ^ (unsigned)((rp & 7) << 4));
            float4 x;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(addr));
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int r = 2 * rp + e;
                const float2 xe = e ? make_float2(x.z, x.w) : make_float2(x.x, x.y);
                if (q < QC) {
                    const float t = g.g[q * D + r];
                    S[0][(q * D + r) % PS] = ffma2(make_float2(t, t), xe, S[0][(q * D + r) % PS]);
                }
                if (q >= 1) {
                    const float t = g.g[(q - 1) * D + r];
                    S[1][((q - 1) * D + r) % PS] = ffma2(make_float2(t, t), xe, S[1][((q - 1) * D + r) % PS]);
                }
            }
        }
    }
}

// What bounds it (profiles/r02_s1_bounds.json, option "s1_diag"): with the filter switched off the TMA ring alone moves a
// 16 Mi-sample chunk in 26 us (0.89 of the measured copy bandwidth) at ring depth 2 or 3; with the loads switched off the
// consumer warps alone need 34 us; the kernel takes 36 us.  Three restructurings of the consumer side were built, validated
// on the whole GPU suite and measured, and none moved that number (the code is in the history):
// two groups of warps sharing a tile with an exchange of partial sums (40 us), ring slots of one 128-byte segment so that a
// segment goes back to the TMA engine after eight phase pairs (37 us), filter warps handing their class sums to combine warps
// through the tile's private rows (36 us, twice the warps per scheduler).
template <int LOGD, int QC, int PS, int MT, int NST>
__global__ void __launch_bounds__((MT / 64 + 1) * 32, 1)
k_xd_tma(const __grid_constant__ XdParams p, const __grid_constant__ XtGeom g, const __grid_constant__ CUtensorMap tm) {
    using Lay = XtLay<LOGD, QC, MT>;
    constexpr int D = Lay::D, NSEG = Lay::NSEG, NR = Lay::NR, REGION = Lay::REGION, STAGE = Lay::STAGE, NW = Lay::NW;
    constexpr int NQH = (QC >> 1) + 1;
    constexpr int NSLOT = NST;                                           // ring depth: a slot is a whole tile (NSEG * 2 boxes)
    constexpr int SLOT = STAGE;
    extern __shared__ __align__(1024) unsigned char xt_smem[];
    // the dynamic window is 1024-byte aligned by the launch (checked on the host side: static shared memory is tiny)
    const unsigned sbase = ((unsigned)__cvta_generic_to_shared(xt_smem) + 1023u) & ~1023u;
    unsigned char* const gbase = xt_smem + (sbase - (unsigned)__cvta_generic_to_shared(xt_smem));
    float2* C = reinterpret_cast<float2*>(gbase + XT_STAGES * STAGE);    // [njobs][PS]  e^{j w_v ((b - s) mod PS)}
    float2* TL = C + (size_t)B200_BATCH * PS;                            // [njobs][16]  e^{j w_v D j}
    float2* PHW = TL + (size_t)B200_BATCH * 16;                          // [NW][njobs][4] per warp and tile: coarse phase
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(PHW + (size_t)NW * B200_BATCH * 4);
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(bars);     // full[s] = bar0 + 8 s, empty[s] = bar0 + 8 (NSLOT + s)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < NSLOT; s++) {
            xt_mbar_init(bar0 + 8 * s, 1);
            xt_mbar_init(bar0 + 8 * (NSLOT + s), NW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    // the producer starts the first tiles right away; the consumers build their tables under those loads
    if (warp < NW) {
        for (int idx = tid; idx < p.njobs * PS; idx += NW * 32) {
            const int v = idx / PS, b = idx - v * PS;
            int a = (b - g.s) % PS;
            if (a < 0) { a += PS; }
            C[idx] = phasor_u64(p.job[v].w * (unsigned long long)a);
        }
        for (int idx = tid; idx < p.njobs * 16; idx += NW * 32) {
            const int v = idx >> 4, j = idx & 15;
            TL[idx] = phasor_u64(p.job[v].w * (unsigned long long)((long long)j * D));
        }
        asm volatile("bar.sync 1, %0;\n" ::"n"(NW * 32) : "memory");
    }

    // a tile is "fast" when every row it reads lies inside the tensor map; the few others (stream history in front of
    // the chunk, the ragged end) get their missing samples patched in by the consumer warps after the TMA boxes landed
    auto tile_fast = [&](long long J0) { return J0 >= 0 && (J0 >> 1) + Lay::NEED <= (long long)g.rows_tma; };

    if (warp == NW) {
        // ---------------- producer ----------------
        if (lane == 0 && !(g.diag & 2)) {
            const unsigned long long pol = xt_policy_evict_first();
            int it = 0;                                                  // slot steps issued so far
            for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
                const long long J0 = g.jmin + (long long)tile * MT;
                const int row0 = (int)(J0 >> 1);
                // every tile comes in through the TMA engine; rows outside the tensor (negative: stream history in front of
                // the chunk; past its last full super-row: the ragged end) arrive zero-filled and are patched by the consumers
                const int sl = it % NSLOT;
                const unsigned ph = (unsigned)((it / NSLOT) & 1);
                it++;
                xt_mbar_wait(bar0 + 8 * (NSLOT + sl), ph ^ 1u);
                const unsigned full = bar0 + 8 * sl;
                xt_mbar_expect(full, (unsigned)SLOT);
                const unsigned dst = sbase + (unsigned)(sl * SLOT);
#pragma unroll
                for (int sg = 0; sg < NSEG; sg++) {
#pragma unroll
                    for (int par = 0; par < 2; par++) {
                        xt_tma_2d(dst + (unsigned)((sg * 2 + par) * REGION), &tm, full, par * 2 * D + sg * 32, row0, pol);
                    }
                }
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    const int jl = warp * 64 + 2 * lane;            // first of this lane's two adjacent outputs (even)
    unsigned aoff[NQH];
#pragma unroll
    for (int qh = 0; qh < NQH; qh++) {
        const unsigned row = (unsigned)((jl >> 1) + qh);
        aoff[qh] = row * 128u + ((row & 7u) << 4);
    }
    int it = 0;                                     // tiles consumed so far (the producer's sequence)
    for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
        const long long J0 = g.jmin + (long long)tile * MT;
        // coarse phase table of this warp's 64 outputs: entry (job, i) = phase at output jl0 + 16 i, drift-centred
        {
            for (int e = lane; e < p.njobs * 4; e += 32) {
                const int vv = e >> 2, i = e & 3;
                const XdJob& Jv = p.job[vv];
                const int a0 = Jv.offset - (Jv.T - 1);
                const long long im0 = (long long)a0 + (J0 + warp * 64 + i * 16 - (long long)g.cj) * D;
                unsigned long long dr = Jv.w * (unsigned long long)PS;
                if (p.pfb_sigma < 0) { dr -= 0x8000000000000000ULL; }
                const long long corr = (long long)dr * (long long)(Jv.T / (2 * PS));
                PHW[(warp * B200_BATCH + vv) * 4 + i] = phasor_u64(Jv.phase0 + Jv.w * (unsigned long long)im0 + (unsigned long long)corr);
            }
        }
        const bool fast = tile_fast(J0);

        // ---- accumulate: D/2 phase pairs x (QC+1) window blocks, 2 outputs x PS classes per lane, slot by slot ----
        float2 S[2][PS];
#pragma unroll
        for (int o = 0; o < 2; o++)
#pragma unroll
            for (int a = 0; a < PS; a++) { S[o][a] = make_float2(0.f, 0.f); }
        const int sl = it % NSLOT;
        const unsigned ph = (unsigned)((it / NSLOT) & 1);
        it++;
        if (!(g.diag & 2)) { xt_mbar_wait(bar0 + 8 * sl, ph); }
        const unsigned sb = sbase + (unsigned)(sl * SLOT);
        if (!fast && !g.diag) {
            // patch what the tensor map does not cover: samples before the chunk (from the carried raw history) and the
            // samples of the last, partial super-row.  A few hundred at most, once or twice per launch.
            const long long ibase = J0 * D + g.org;
            unsigned char* stg = gbase + sl * SLOT;
            constexpr int NSAMP = (MT + QC + 1) * D;
            const long long tensor_end = (long long)g.org + (long long)g.rows_tma * (2 * D);
            // the tensor's row 0 starts at sample g.org: everything of this tile in front of it was zero-filled -- stream
            // history (index < 0) AND the first g.org samples of the chunk
            const long long front = (long long)g.org - ibase;
            const int head = (int)(front > 0 ? (front < NSAMP ? front : NSAMP) : 0);
            long long t0 = tensor_end - ibase, t1 = (long long)p.count - ibase;
            if (t0 < head) { t0 = head; }
            if (t1 > NSAMP) { t1 = NSAMP; }
            auto patch = [&](int idx) {
                const float2 v = load_x<FMT_CF32>(p, ibase + idx);
                const int j = idx >> LOGD, r = idx & (D - 1), row = j >> 1;
                const int off = ((r >> 4) * 2 + (j & 1)) * REGION + row * 128 + ((((r >> 1) & 7) ^ (row & 7)) << 4) + (r & 1) * 8;
                *reinterpret_cast<float2*>(stg + off) = v;
            };
            for (int idx = tid; idx < head; idx += NW * 32) { patch(idx); }
            for (long long idx = t0 + tid; idx < t1; idx += NW * 32) { patch((int)idx); }
            asm volatile("bar.sync 1, %0;\n" ::"n"(NW * 32) : "memory");
        }
        if (!(g.diag & 1)) {
            xt_accumulate<LOGD, QC, PS, REGION, 0, 8, NQH>(S, g, sb, aoff);
            xt_accumulate<LOGD, QC, PS, REGION, 8, 16, NQH>(S, g, sb + 2 * REGION, aoff);
            if constexpr (NSEG == 4) {
                xt_accumulate<LOGD, QC, PS, REGION, 16, 24, NQH>(S, g, sb + 4 * REGION, aoff);
                xt_accumulate<LOGD, QC, PS, REGION, 24, 32, NQH>(S, g, sb + 6 * REGION, aoff);
            }
        }
        static_assert(NSEG == 2 || NSEG == 4, "D = 32 or 64");
        // the tile is consumed: hand it back to the producer before the (register-only) combination
        __syncwarp();
        if (lane == 0 && !(g.diag & 2)) { xt_mbar_arrive(bar0 + 8 * (NSLOT + sl)); }

        if (g.diag & 1) { continue; }
        // ---- combine per slot (a VFO, or a +f / -f pair sharing A = sum cos*S and B = sum sin*S), rotate, store ----
        const long long m0 = J0 + jl - (long long)g.cj;
        const int nout = p.job[0].n_out;                       // every job of a filter-bank launch has the same length
        const bool ok0 = m0 >= 0 && m0 < nout, ok1 = (m0 + 1) >= 0 && (m0 + 1) < nout;
        const int ih = (2 * lane) >> 4, il = (2 * lane) & 15;
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
                const float2 pc = PHW[(warp * B200_BATCH + ja) * 4 + ih];
                const float2 p0 = cmulf(pc, TL[ja * 16 + il]), p1 = cmulf(pc, TL[ja * 16 + il + 1]);
                float2* out = p.job[ja].out + m0;
                if (ok0) { out[0] = cmulf(make_float2(A0.x - B0.y, A0.y + B0.x), p0); }
                if (ok1) { out[1] = cmulf(make_float2(A1.x - B1.y, A1.y + B1.x), p1); }
            }
            if (jb >= 0) {
                const float2 pc = PHW[(warp * B200_BATCH + jb) * 4 + ih];
                const float2 p0 = cmulf(pc, TL[jb * 16 + il]), p1 = cmulf(pc, TL[jb * 16 + il + 1]);
                float2* out = p.job[jb].out + m0;
                if (ok0) { out[0] = cmulf(make_float2(A0.x + B0.y, A0.y - B0.x), p0); }      // conjugate coefficients
                if (ok1) { out[1] = cmulf(make_float2(A1.x + B1.y, A1.y - B1.x), p1); }
            }
        }
        __syncwarp();                                          // PHW is rewritten for the next tile
    }
}
