// sdrplusplus_b200/csrc/fused_tail.cuh -- k_tail_fused: all FIR-like stages after stage 1 of a VFO in one launch
// (included by kernels.cu; descriptors and the shared host/device range arithmetic are in kernels.cuh).
//
// Replaces, for one VFO and one chunk, the per-stage launches of
//   DecimatingFIR stages 2..k        (power_decimator.h:58-65, decimating_fir.h:45-68)
//   PolyphaseResampler               (polyphase_resampler.h:69-99)
//   FIR<complex_t,float>             (rx_vfo.h:28-31 channel filter, fir.h:62-83)
//   Quadrature                       (quadrature.h:39-46)
//   FIR<float,float> + LRToStereo    (broadcast_fm.h:45, fm.h:123)
// Shared-memory arena: stage s's input lives at float offset st[s].buf in a PHASE-MAJOR layout -- sample i of the
// range sits in row (i - o) % D, column (i - o) / D, D the stage's decimation -- so that a thread producing FT_R
// consecutive outputs slides one register window along a row per tap phase: one LDS per FT_R packed FMAs.
// Taps are stored phase-major too (row r = taps[q*D + r]) and read as warp-wide broadcasts.
#pragma once

struct FtLay { int o, D, pitch; };

__device__ __forceinline__ FtLay ft_make_lay(const FtStage& S, int lo) {
    FtLay l;
    l.o = ft_origin(S, lo);
    l.D = ft_rows(S);
    l.pitch = S.pitch;
    return l;
}
__device__ __forceinline__ int ft_lidx(const FtLay& l, int i) {
    const int rel = i - l.o;
    if (l.D == 1) { return rel; }
    const int j = rel / l.D;
    return (rel - j * l.D) * l.pitch + j;
}

__device__ __forceinline__ float2 ft_fma(float h, float2 x, float2 a) { return ffma2(make_float2(h, h), x, a); }
__device__ __forceinline__ float ft_fma(float h, float x, float a) { return fmaf(h, x, a); }
__device__ __forceinline__ void ft_zero(float2& a) { a = make_float2(0.0f, 0.0f); }
__device__ __forceinline__ void ft_zero(float& a) { a = 0.0f; }

// one block of N taps: all loads first (taps + the window refills), then N * FT_R independent multiply-adds.
// The last block of a run needs only the refills its own later taps read (N - 1 of them).
template <typename T, int N, bool LAST>
__device__ __forceinline__ void ft_blk(T (&acc)[FT_R], T (&w)[FT_R], const T* __restrict__ X, const float* __restrict__ taps) {
    constexpr int NR = LAST ? N - 1 : N;
    float h[N];
    T nw[NR > 0 ? NR : 1];
#pragma unroll
    for (int k = 0; k < N; k++) { h[k] = taps[k]; }
#pragma unroll
    for (int k = 0; k < NR; k++) { nw[k] = X[k + FT_R]; }
#pragma unroll
    for (int k = 0; k < N; k++) {
#pragma unroll
        for (int i = 0; i < FT_R; i++) { acc[i] = ft_fma(h[k], w[(k + i) % FT_R], acc[i]); }
        if (k < NR) { w[k] = nw[k]; }
    }
}
// acc[i] += sum_{q < Q} taps[q] * X[i + q],  Q >= 1
template <typename T>
__device__ __forceinline__ void ft_firx(T (&acc)[FT_R], const T* __restrict__ X, const float* __restrict__ taps, int Q) {
    T w[FT_R];
#pragma unroll
    for (int i = 0; i < FT_R; i++) { w[i] = X[i]; }
    int q0 = 0;
    for (; q0 + FT_R < Q; q0 += FT_R) { ft_blk<T, FT_R, false>(acc, w, X + q0, taps + q0); }
    X += q0; taps += q0;
    switch (Q - q0) {
    case 1: ft_blk<T, 1, true>(acc, w, X, taps); break;
    case 2: ft_blk<T, 2, true>(acc, w, X, taps); break;
    case 3: ft_blk<T, 3, true>(acc, w, X, taps); break;
    case 4: ft_blk<T, 4, true>(acc, w, X, taps); break;
    case 5: ft_blk<T, 5, true>(acc, w, X, taps); break;
    case 6: ft_blk<T, 6, true>(acc, w, X, taps); break;
    case 7: ft_blk<T, 7, true>(acc, w, X, taps); break;
    case 8: ft_blk<T, 8, true>(acc, w, X, taps); break;
    case 9: ft_blk<T, 9, true>(acc, w, X, taps); break;
    default: break;
    }
}
// A unit (FT_R outputs) can be shared by S = 1, 2, 4 or 8 adjacent lanes, each taking every S-th (row, tap segment)
// piece of the sum; the partial sums meet in a butterfly.  Picks S so that a phase with few units still uses
// the whole CTA: the phases of a slab are serial, their length is what the slab costs.
__device__ __forceinline__ int ft_pick_split(int units, int nthreads, int max_pieces) {
    int S = 1;
    while (S < 8 && units * S < nthreads && 2 * S <= max_pieces) { S <<= 1; }
    return S;
}
__device__ __forceinline__ void ft_reduce(float2 (&acc)[FT_R], int S) {
    for (int d = 1; d < S; d <<= 1) {
#pragma unroll
        for (int i = 0; i < FT_R; i++) {
            acc[i].x += __shfl_xor_sync(0xffffffffu, acc[i].x, d);
            acc[i].y += __shfl_xor_sync(0xffffffffu, acc[i].y, d);
        }
    }
}
__device__ __forceinline__ void ft_reduce(float (&acc)[FT_R], int S) {
    for (int d = 1; d < S; d <<= 1) {
#pragma unroll
        for (int i = 0; i < FT_R; i++) { acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], d); }
    }
}

// where a stage's outputs go: the next stage's shared-memory buffer, or global memory for the last stage
struct FtDst {
    float* g;           // non-null: global
    float* sp;          // shared buffer base
    FtLay lay;
    int dup;
    __device__ __forceinline__ void put2(int m, float2 v) const {
        if (g) { reinterpret_cast<float2*>(g)[m] = v; }
        else { reinterpret_cast<float2*>(sp)[ft_lidx(lay, m)] = v; }
    }
    __device__ __forceinline__ void put1(int m, float v) const {
        if (dup) { put2(m, make_float2(v, v)); }
        else if (g) { g[m] = v; }
        else { sp[ft_lidx(lay, m)] = v; }
    }
    // FT_R consecutive outputs m .. m+FT_R-1 (those below mend): one division, then a (row, column) cursor
    __device__ __forceinline__ void put2_run(int m, int mend, const float2 (&v)[FT_R]) const {
        if (g) {
#pragma unroll
            for (int i = 0; i < FT_R; i++) { if (m + i < mend) { reinterpret_cast<float2*>(g)[m + i] = v[i]; } }
            return;
        }
        float2* b = reinterpret_cast<float2*>(sp);
        if (lay.D == 1) {
#pragma unroll
            for (int i = 0; i < FT_R; i++) { if (m + i < mend) { b[m - lay.o + i] = v[i]; } }
            return;
        }
        const int rel = m - lay.o;
        int j = rel / lay.D, r = rel - j * lay.D;
        int idx = r * lay.pitch + j;
#pragma unroll
        for (int i = 0; i < FT_R; i++) {
            if (m + i < mend) { b[idx] = v[i]; }
            r++; idx += lay.pitch;
            if (r == lay.D) { r = 0; idx -= lay.D * lay.pitch - 1; }
        }
    }
};

template <int NT>
__global__ void __launch_bounds__(NT, 2) k_tail_fused(const __grid_constant__ FtParams p) {
    extern __shared__ __align__(16) float sm[];
    __shared__ int s_lo[FT_MAXST + 1], s_hi[FT_MAXST + 1];
    const FtJob& J = p.job[blockIdx.y];
    const int slab = blockIdx.x;
    if (slab >= J.slabs) { return; }
    const int nst = J.nst;
    const int tid = threadIdx.x;
    const bool last = (slab == J.slabs - 1);
    if (tid == 0) {
        int lo[FT_MAXST + 1], hi[FT_MAXST + 1];
        ft_ranges(J, slab, lo, hi);
        for (int s = 0; s <= nst; s++) { s_lo[s] = lo[s]; s_hi[s] = hi[s]; }
    }
    // ---- taps -> shared memory (already phase-major in global memory: Stage::taps_pm) ----
    for (int s = 0; s < nst; s++) {
        const FtStage& S = J.st[s];
        float4* tp = reinterpret_cast<float4*>(sm + S.tap_off);
        const float4* __restrict__ src = reinterpret_cast<const float4*>(S.taps);
        for (int k = tid; k < (S.ntap_f >> 2); k += NT) { tp[k] = __ldg(src + k); }
    }
    __syncthreads();

    for (int s = 0; s < nst; s++) {
        const FtStage& S = J.st[s];
        const bool fin = (s + 1 == nst);
        const int mlo = max(s_lo[s + 1], 0), mhi = s_hi[s + 1];
        FtDst dst;
        dst.g = fin ? J.out : nullptr;
        dst.dup = S.dup;
        dst.sp = sm;
        dst.lay.o = 0; dst.lay.D = 1; dst.lay.pitch = 0;
        if (!fin) {
            const FtStage& N = J.st[s + 1];
            dst.sp = sm + N.buf;
            dst.lay = ft_make_lay(N, s_lo[s + 1]);
            // history part of the next stage's input
            const int hend = min(0, s_hi[s + 1]);
            if (N.es == 2) {
                for (int i = s_lo[s + 1] + tid; i < hend; i += NT) {
                    reinterpret_cast<float2*>(dst.sp)[ft_lidx(dst.lay, i)] = __ldg(reinterpret_cast<const float2*>(N.hist_rd) + (i + N.hist));
                }
            }
            else {
                for (int i = s_lo[s + 1] + tid; i < hend; i += NT) { dst.sp[ft_lidx(dst.lay, i)] = __ldg(N.hist_rd + (i + N.hist)); }
            }
        }
        // stage 0 streams its input from global memory through the staging buffer, sub-tile by sub-tile
        int tl = (s == 0) ? J.OT0 : (mhi - mlo);
        if (tl < 1) { tl = 1; }
        float* xb = sm + S.buf;
        for (int t0 = mlo; t0 < mhi; t0 += tl) {
            const int t1 = min(mhi, t0 + tl);
            FtLay ls;
            if (s == 0) {
                int ilo, ihi;
                ft_need_in(S, t0, t1, ilo, ihi);
                ls = ft_make_lay(S, ilo);
                const int n = ihi - ls.o;
                // idx = i - o walks the rows round-robin: row = idx % D, column = idx / D, advanced without dividing
                const int D = ls.D;
                int col = tid / D, row = tid - col * D;
                const int dcol = NT / D, drow = NT - dcol * D;
                if (S.es == 2) {
                    const float2* __restrict__ src = reinterpret_cast<const float2*>(J.src) + (ls.o + S.hist);
                    const int imin = -S.hist - ls.o;
                    for (int idx = tid; idx < n; idx += NT) {
                        reinterpret_cast<float2*>(xb)[row * ls.pitch + col] = (idx >= imin) ? __ldg(src + idx) : make_float2(0.0f, 0.0f);
                        row += drow; col += dcol;
                        if (row >= D) { row -= D; col++; }
                    }
                }
                else {
                    const float* __restrict__ src = J.src + (ls.o + S.hist);
                    const int imin = -S.hist - ls.o;
                    for (int idx = tid; idx < n; idx += NT) {
                        xb[row * ls.pitch + col] = (idx >= imin) ? __ldg(src + idx) : 0.0f;
                        row += drow; col += dcol;
                        if (row >= D) { row -= D; col++; }
                    }
                }
                __syncthreads();
            }
            else { ls = ft_make_lay(S, s_lo[s]); }

            switch (S.kind) {
            case FT_FIRC: {
                const float2* X = reinterpret_cast<const float2*>(xb);
                const float* tp = sm + S.tap_off;
                const int D = S.D, T = S.T;
                const int m_o = (ls.o - (S.off - (T - 1))) / D;       // exact for D > 1; D == 1: o - c
                const int qf = T / D, rem = T - qf * D;
                const int units = (t1 - t0 + FT_R - 1) / FT_R;
                const int qmax = qf + (rem ? 1 : 0);
                const int SP = ft_pick_split(units, NT, D * max(1, qmax / 8));
                const int nseg = (SP + D - 1) / D;                     // tap segments per row
                const int qs = (qmax + nseg - 1) / nseg;
                const int items = units * SP;
                for (int w0 = 0; w0 < items; w0 += NT) {
                    const int w = w0 + tid;
                    const bool act = w < items;
                    const int u = w >> (31 - __clz(SP)), sidx = w & (SP - 1);
                    const int m = t0 + u * FT_R;
                    float2 acc[FT_R];
#pragma unroll
                    for (int i = 0; i < FT_R; i++) { ft_zero(acc[i]); }
                    if (act) {
                        const int pos = m - m_o;
                        for (int seg = sidx; seg < D * nseg; seg += SP) {
                            const int r = seg / nseg, part = seg - r * nseg;
                            const int qa = part * qs, qb = min(qf + (r < rem ? 1 : 0), qa + qs);
                            if (qb > qa) { ft_firx<float2>(acc, X + r * ls.pitch + pos + qa, tp + r * S.qpitch + qa, qb - qa); }
                        }
                    }
                    ft_reduce(acc, SP);
                    if (act && sidx == 0) { dst.put2_run(m, t1, acc); }
                }
                break;
            }
            case FT_FIRR: {
                const float* X = xb;
                const float* tp = sm + S.tap_off;
                const int m_o = ls.o - (S.off - (S.T - 1));
                const int units = (t1 - t0 + FT_R - 1) / FT_R;
                const int SP = ft_pick_split(units, NT, max(1, S.T / 8));
                const int qs = (S.T + SP - 1) / SP;
                const int items = units * SP;
                for (int w0 = 0; w0 < items; w0 += NT) {
                    const int w = w0 + tid;
                    const bool act = w < items;
                    const int u = w >> (31 - __clz(SP)), sidx = w & (SP - 1);
                    const int m = t0 + u * FT_R;
                    float acc[FT_R];
#pragma unroll
                    for (int i = 0; i < FT_R; i++) { ft_zero(acc[i]); }
                    if (act) {
                        const int qa = sidx * qs, qb = min(S.T, qa + qs);
                        if (qb > qa) { ft_firx<float>(acc, X + (m - m_o) + qa, tp + qa, qb - qa); }
                    }
                    ft_reduce(acc, SP);
                    if (act && sidx == 0) {
#pragma unroll
                        for (int i = 0; i < FT_R; i++) {
                            if (m + i < t1) { dst.put1(m + i, acc[i]); }
                        }
                    }
                }
                break;
            }
            case FT_POLY: {
                const float2* X = reinterpret_cast<const float2*>(xb);
                const float* tp = sm + S.tap_off;
                const int L = S.L, M = S.D, tpp = S.T;
                const int qf = tpp / M, qrem = tpp - qf * M;
                const int per_class = (t1 - t0 + L - 1) / L;
                const int gmax = (per_class + FT_R - 1) / FT_R;
                const int units = L * gmax;
                const int SP = ft_pick_split(units, NT, M);            // rows of the phase-major input are the pieces
                const int items = units * SP;
                for (int w0 = 0; w0 < items; w0 += NT) {
                    const int w = w0 + tid;
                    const int u = w >> (31 - __clz(SP)), sidx = w & (SP - 1);
                    const int c = u / gmax, g = u - c * gmax;
                    const int m_c = t0 + ft_posmod((long long)c - t0, L);          // first output of this class in the tile
                    const int cnt = (m_c < t1) ? (t1 - m_c + L - 1) / L : 0;
                    const int jj0 = g * FT_R;
                    const bool act = (w < items) && (jj0 < cnt);
                    float2 acc[FT_R];
#pragma unroll
                    for (int i = 0; i < FT_R; i++) { ft_zero(acc[i]); }
                    if (act) {
                        const long long t = (long long)S.phase + (long long)m_c * M;
                        const int ph = (int)(t % L);
                        const int e = S.off + (int)(t / L) - (tpp - 1) - ls.o;     // >= 0
                        const int a = e / M, bb = e - a * M;
                        for (int r = sidx; r < M; r += SP) {
                            int rp = r - bb, cc = 0;
                            if (rp < 0) { rp += M; cc = 1; }
                            const int Q = qf + (rp < qrem ? 1 : 0);
                            if (Q < 1) { continue; }
                            ft_firx<float2>(acc, X + r * ls.pitch + a + cc + jj0, tp + (ph * M + rp) * S.qpitch, Q);
                        }
                    }
                    ft_reduce(acc, SP);
                    if (act && sidx == 0) {
#pragma unroll
                        for (int i = 0; i < FT_R; i++) {
                            if (jj0 + i < cnt) { dst.put2(m_c + (jj0 + i) * L, acc[i]); }
                        }
                    }
                }
                break;
            }
            case FT_QUAD: {
                const float2* X = reinterpret_cast<const float2*>(xb);
                for (int base = t0; base < t1; base += NT) {
                    const int m = base + tid;
                    const bool act = m < t1;
                    float2 cs = act ? X[m - ls.o] : make_float2(0.0f, 0.0f);
                    const float cur = atan2f(cs.y, cs.x);
                    float prev = __shfl_up_sync(0xffffffffu, cur, 1);
                    if ((tid & 31) == 0 && act) {
                        const float2 q = X[m - 1 - ls.o];
                        prev = atan2f(q.y, q.x);
                    }
                    if (act) {
                        float diff = __fsub_rn(cur, prev);
                        if (diff > FL_M_PI_REF) { diff = __fsub_rn(diff, 2.0f * FL_M_PI_REF); }
                        else if (diff <= -FL_M_PI_REF) { diff = __fadd_rn(diff, 2.0f * FL_M_PI_REF); }
                        dst.put1(m, __fmul_rn(diff, S.scale));
                    }
                }
                break;
            }
            default: {   // FT_M2S
                for (int m = t0 + tid; m < t1; m += NT) { dst.put1(m, xb[m - ls.o]); }
                break;
            }
            }
            if (s == 0) { __syncthreads(); }
        }
        __syncthreads();
        // ---- the last slab hands the next stage's final `hist` inputs to the next chunk ----
        if (!fin && last) {
            const FtStage& N = J.st[s + 1];
            const int i0 = N.n_in - N.hist;
            if (N.es == 2) {
                for (int j = tid; j < N.hist; j += NT) {
                    reinterpret_cast<float2*>(N.hist_wr)[j] = reinterpret_cast<const float2*>(dst.sp)[ft_lidx(dst.lay, i0 + j)];
                }
            }
            else {
                for (int j = tid; j < N.hist; j += NT) { N.hist_wr[j] = dst.sp[ft_lidx(dst.lay, i0 + j)]; }
            }
        }
    }
}
