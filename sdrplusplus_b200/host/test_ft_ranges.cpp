// Host-only check of the fused tail's range arithmetic (csrc/kernels.cuh: ft_need_in / ft_ranges), the code the CUDA
// kernel and the scheduler share.  For random WFM-like stage lists, chunk sizes, decimation offsets and slab sizes:
//   * every slab's input range of every stage stays inside [-hist, n_in)
//   * the slabs' final ranges tile [0, n_out) exactly
//   * what stage s produces for a slab covers what stage s+1 reads of the current chunk
//   * the last slab covers the final `hist` inputs of every stage (next chunk's history)
// Exit code 0 = all invariants hold.  Built by __graft_entry__.build(), run by tests/test_abi.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include "../csrc/kernels.cuh"

static unsigned long long rng = 0x9E3779B97F4A7C15ULL;
static int rnd(int n) { rng = rng * 6364136223846793005ULL + 1442695040888963407ULL; return (int)((rng >> 33) % (unsigned long long)n); }

static int out_count(const FtStage& s, int n, int& next_off, int& next_phase) {   // mirrors FirCStage / PolyStage::plan
    if (s.kind == FT_FIRC) {
        int no = (s.off < n) ? (n - s.off + s.D - 1) / s.D : 0;
        next_off = s.off + no * s.D - n;
        return no;
    }
    if (s.kind == FT_POLY) {
        long long avail = ((long long)n - s.off) * s.L - s.phase;
        long long no = avail > 0 ? (avail + s.D - 1) / s.D : 0;
        long long tend = (long long)s.phase + no * s.D;
        next_off = (int)((long long)s.off + tend / s.L - n);
        next_phase = (int)(tend % s.L);
        return (int)no;
    }
    return n;
}

int main() {
    int fails = 0;
    for (int trial = 0; trial < 4000; trial++) {
        FtJob J;
        memset(&J, 0, sizeof(J));
        // decimating FIRs, a polyphase resampler, a channel FIR, the discriminator, an audio FIR
        int nst = 0;
        const int nd = rnd(3);
        for (int i = 0; i < nd; i++) { FtStage& s = J.st[nst++]; s.kind = FT_FIRC; s.D = 2 << rnd(2); s.T = 5 + rnd(70); s.hist = s.T - 1; s.es = 2; s.off = rnd(s.D); }
        if (rnd(2)) { FtStage& s = J.st[nst++]; s.kind = FT_POLY; s.L = 1 + rnd(24); s.D = 1 + rnd(30); s.T = 3 + rnd(120); s.hist = s.T - 1; s.es = 2; s.phase = rnd(s.L); s.off = rnd(3); }
        { FtStage& s = J.st[nst++]; s.kind = FT_FIRC; s.D = 1; s.T = 1 + rnd(130); s.hist = s.T - 1; s.es = 2; s.off = 0; }
        { FtStage& s = J.st[nst++]; s.kind = FT_QUAD; s.D = 1; s.T = 1; s.L = 1; s.hist = 1; s.es = 2; }
        if (rnd(2)) { FtStage& s = J.st[nst++]; s.kind = FT_FIRR; s.D = 1; s.T = 1 + rnd(240); s.hist = s.T - 1; s.es = 1; }
        else { FtStage& s = J.st[nst++]; s.kind = FT_M2S; s.D = 1; s.T = 1; s.hist = 0; s.es = 1; }
        for (int i = 0; i < nst; i++) { if (J.st[i].L == 0) { J.st[i].L = 1; } }
        J.nst = nst;
        int n = rnd(4) == 0 ? rnd(40) : 200 + rnd(60000);
        for (int i = 0; i < nst; i++) {
            int no = 0, np = 0;
            J.st[i].n_in = n;
            J.st[i].n_out = out_count(J.st[i], n, no, np);
            n = J.st[i].n_out;
        }
        J.OB = 36 + rnd(1500);
        const int n_last = J.st[nst - 1].n_out;
        J.slabs = std::max(1, (n_last + J.OB - 1) / J.OB);
        int covered = 0;
        for (int slab = 0; slab < J.slabs; slab++) {
            int lo[FT_MAXST + 1], hi[FT_MAXST + 1];
            ft_ranges(J, slab, lo, hi);
            if (lo[nst] != std::min(covered, n_last) || hi[nst] < lo[nst]) { fails++; fprintf(stderr, "trial %d slab %d: final range [%d,%d) after %d\n", trial, slab, lo[nst], hi[nst], covered); }
            covered = hi[nst];
            for (int s = 0; s < nst; s++) {
                const FtStage& S = J.st[s];
                if (hi[s] > lo[s] && (lo[s] < -S.hist || hi[s] > S.n_in)) { fails++; fprintf(stderr, "trial %d slab %d stage %d: [%d,%d) outside [-%d,%d)\n", trial, slab, s, lo[s], hi[s], S.hist, S.n_in); }
                // what this stage computes = the non-negative part of the next stage's range: it must not exceed n_out
                if (hi[s + 1] > S.n_out) { fails++; fprintf(stderr, "trial %d slab %d stage %d produces past n_out\n", trial, slab, s); }
                // and its input need must be met by [lo[s], hi[s])
                const int pl = std::max(lo[s + 1], 0), ph = hi[s + 1];
                if (ph > pl) {
                    int ilo, ihi;
                    ft_need_in(S, pl, ph, ilo, ihi);
                    if (ilo < lo[s] || ihi > hi[s]) { fails++; fprintf(stderr, "trial %d slab %d stage %d: need [%d,%d) not in [%d,%d)\n", trial, slab, s, ilo, ihi, lo[s], hi[s]); }
                }
                if (slab == J.slabs - 1 && s > 0 && (lo[s] > S.n_in - S.hist || hi[s] != S.n_in)) { fails++; fprintf(stderr, "trial %d: last slab misses the history of stage %d\n", trial, s); }
            }
            if (fails > 20) { return 1; }
        }
        if (covered != n_last) { fails++; fprintf(stderr, "trial %d: slabs cover %d of %d\n", trial, covered, n_last); }
    }
    printf("ft_ranges: %s\n", fails ? "FAILED" : "ok");
    return fails ? 1 : 0;
}
