/*
 * oracle/sdrpp_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of SDR++'s streaming DSP hot path,
 * function by function, each citing the reference file:line it follows.  The leaf
 * arithmetic is oracle/volk_generic.h + oracle/offt.h (VOLK and FFTW are external,
 * un-pinned and absent: PARITY UNPINNED at that boundary, see sdrpp_oracle.h).
 * This file is required to be BIT-IDENTICAL to oracle/_ref/libsdrpp_ref.so (the
 * reference's own headers over the same leaf layer) when both are built with
 * -O2 -ffp-contract=off; tests/test_oracle_vs_ref.py and the fixtures under
 * tests/golden/ enforce that.
 *
 * Nothing in the product (sdrplusplus_b200/, include/) may link or call this file.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sdrpp_oracle.h"
#include "volk_generic.h"
#include "offt.h"

#define DB_M_PI 3.14159265358979323846 /* core/src/dsp/math/constants.h:3 */
#define FL_M_PI 3.1415926535f          /* core/src/dsp/math/constants.h:4 */
#define STREAM_BUFFER_SIZE 1000000     /* core/src/dsp/stream.h:9 */
#define DELAY_EXTRA 64000              /* core/src/dsp/filter/fir.h:24 */

typedef ovk_cf32 cf32;

const char* orc_impl(void) { return "restatement"; }

/* ------------------------------------------------------------------ */
/* host-side design                                                    */
/* ------------------------------------------------------------------ */

/* math::hzToRads  (core/src/dsp/math/hz_to_rads.h:6-8) */
static double hz_to_rads(double freq, double samplerate) { return 2.0 * DB_M_PI * (freq / samplerate); }

/* math::sinc  (core/src/dsp/math/sinc.h:5-7) */
static double sinc_d(double x) { return (x == 0.0) ? 1.0 : (sin(x) / x); }

/* window::cosine  (core/src/dsp/window/cosine.h:7-15) */
static double win_cosine(double n, double N, const double* coefs, int coefCount) {
    double win = 0.0;
    double sign = 1.0;
    for (int i = 0; i < coefCount; i++) {
        win += sign * coefs[i] * cos((double)i * 2.0 * DB_M_PI * n / N);
        sign = -sign;
    }
    return win;
}
/* window::nuttall (window/nuttall.h:5-8), window::blackman (window/blackman.h:5-8) */
static double win_nuttall(double n, double N) {
    const double coefs[] = { 0.355768, 0.487396, 0.144232, 0.012604 };
    return win_cosine(n, N, coefs, 4);
}
static double win_blackman(double n, double N) {
    const double coefs[] = { 0.42, 0.5, 0.08 };
    return win_cosine(n, N, coefs, 3);
}

double orc_window(int type, double n, double N) {
    if (type == 1) { return win_blackman(n, N); }
    if (type == 2) { return win_nuttall(n, N); }
    return 1.0;
}

/* taps::estimateTapCount  (core/src/dsp/taps/estimate_tap_count.h:4-6): double -> int truncation */
int orc_estimate_tap_count(double transWidth, double samplerate) { return (int)(3.8 * samplerate / transWidth); }

/* taps::windowedSinc<float>(count, omega, nuttall)  (core/src/dsp/taps/windowed_sinc.h:9-29) */
static float* windowed_sinc_f(int count, double omega) {
    float* taps = (float*)malloc(sizeof(float) * (size_t)(count > 0 ? count : 1));
    double half = (double)count / 2.0;
    double corr = 1.0 * omega / DB_M_PI;
    for (int i = 0; i < count; i++) {
        double t = (double)i - half + 0.5;
        taps[i] = (float)(sinc_d(t * omega) * win_nuttall(t - half, count) * corr);
    }
    return taps;
}

/* taps::lowPass  (core/src/dsp/taps/low_pass.h:7-11) */
static float* lowpass_taps(double cutoff, double transWidth, double sampleRate, int odd, int* countOut) {
    int count = orc_estimate_tap_count(transWidth, sampleRate);
    if (odd && !(count % 2)) { count++; }
    *countOut = count;
    return windowed_sinc_f(count, hz_to_rads(cutoff, sampleRate));
}

int orc_lowpass(double cutoff, double transWidth, double samplerate, int odd, float* out, int cap) {
    int n;
    float* t = lowpass_taps(cutoff, transWidth, samplerate, odd, &n);
    if (out) { memcpy(out, t, sizeof(float) * (size_t)(n < cap ? n : cap)); }
    free(t);
    return n;
}

/* taps::highPass  (core/src/dsp/taps/high_pass.h:7-14): windowedSinc at (fs/2 - cutoff) with the window
 * nuttall(n,N) * ((int)round(n) % 2 ? -1.0f : 1.0f)  (double * float -> double) */
int orc_highpass(double cutoff, double transWidth, double sampleRate, int odd, float* out, int cap) {
    int count = orc_estimate_tap_count(transWidth, sampleRate);
    if (odd && !(count % 2)) { count++; }
    const double omega = hz_to_rads((sampleRate / 2.0) - cutoff, sampleRate);
    const double half = (double)count / 2.0;
    const double corr = 1.0 * omega / DB_M_PI;
    for (int i = 0; i < count && i < cap; i++) {
        double t = (double)i - half + 0.5;
        double n = t - half;
        double w = win_nuttall(n, count) * ((((int)round(n)) % 2) ? -1.0f : 1.0f);
        if (out) { out[i] = (float)(sinc_d(t * omega) * w * corr); }
    }
    return count;
}

/* taps::bandPass<complex_t>  (core/src/dsp/taps/band_pass.h:11-26) with
 * windowedSinc<complex_t> (windowed_sinc.h:22-25): cplx{(float)sinc,0} * window(..) * corr, where the
 * window is a complex_t (phasor * double -> float multiply, types.h:12-14) and the final "* corr" is the
 * complex_t::operator*(double) overload (float multiply by (float)corr). */
int orc_bandpass_c(double bandStart, double bandStop, double transWidth, double sampleRate, int odd, float* out, int cap) {
    float offsetOmega = (float)hz_to_rads((bandStart + bandStop) / 2.0, sampleRate);
    int count = orc_estimate_tap_count(transWidth, sampleRate);
    if (odd && !(count % 2)) { count++; }
    double omega = hz_to_rads((bandStop - bandStart) / 2.0, sampleRate);
    double half = (double)count / 2.0;
    double corr = 1.0 * omega / DB_M_PI;
    for (int i = 0; i < count && i < cap; i++) {
        double t = (double)i - half + 0.5;
        double n = t - half;
        float x = -offsetOmega * (float)n;
        cf32 ph = { cosf(x), sinf(x) };               /* math::phasor (math/phasor.h:6-9) */
        double wn = win_nuttall(n, count);
        cf32 w = { ph.re * (float)wn, ph.im * (float)wn }; /* complex_t * double */
        cf32 c = { (float)sinc_d(t * omega), 0.0f };
        cf32 p;                                       /* complex_t * complex_t (types.h:24-26) */
        p.re = (c.re * w.re) - (c.im * w.im);
        p.im = (c.im * w.re) + (c.re * w.im);
        if (out) {
            out[2 * i] = p.re * (float)corr;
            out[2 * i + 1] = p.im * (float)corr;
        }
    }
    return count;
}

/* ---- decimation plans: numeric tables of decim/plans.h + decim/taps/*.h, loaded from the data file
 *      written by tools/extract_decim_plans.py ---- */
#define MAX_PLANS 16
#define MAX_STAGES 8
typedef struct { int decim, ntaps; float* taps; } plan_stage;
typedef struct { int ratio, nstages; plan_stage st[MAX_STAGES]; } decim_plan;
static decim_plan g_plans[MAX_PLANS];
static int g_nplans = -1;

static void load_plans(void) {
    if (g_nplans >= 0) { return; }
    g_nplans = 0;
    char path[4096];
    const char* env = getenv("SDRPP_DECIM_PLANS");
    if (env) {
        snprintf(path, sizeof(path), "%s", env);
    }
    else {
        Dl_info info;
        if (!dladdr((void*)&load_plans, &info) || !info.dli_fname) { return; }
        snprintf(path, sizeof(path), "%s", info.dli_fname);
        char* slash = strrchr(path, '/');
        if (!slash) { return; }
        *slash = 0;
        strncat(path, "/../sdrplusplus_b200/data/decim_plans.bin", sizeof(path) - strlen(path) - 1);
    }
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "oracle: cannot open %s\n", path); return; }
    char magic[8];
    int32_t n = 0;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "SDRPPDP1", 8) || fread(&n, 4, 1, f) != 1) { fclose(f); return; }
    for (int p = 0; p < n && p < MAX_PLANS; p++) {
        int32_t hdr[2];
        if (fread(hdr, 4, 2, f) != 2) { break; }
        g_plans[p].ratio = hdr[0];
        g_plans[p].nstages = hdr[1];
        for (int s = 0; s < hdr[1]; s++) {
            int32_t sh[2];
            if (fread(sh, 4, 2, f) != 2) { break; }
            g_plans[p].st[s].decim = sh[0];
            g_plans[p].st[s].ntaps = sh[1];
            g_plans[p].st[s].taps = (float*)malloc(sizeof(float) * (size_t)sh[1]);
            if (fread(g_plans[p].st[s].taps, 4, (size_t)sh[1], f) != (size_t)sh[1]) { break; }
        }
        g_nplans++;
    }
    fclose(f);
}

static const decim_plan* find_plan(int ratio) {
    load_plans();
    for (int p = 0; p < g_nplans; p++) {
        if (g_plans[p].ratio == ratio) { return &g_plans[p]; }
    }
    return NULL;
}

int orc_decim_plan(int ratio, int* decims, int* tapcounts, int cap) {
    const decim_plan* p = find_plan(ratio);
    if (!p) { return 0; }
    for (int i = 0; i < p->nstages && i < cap; i++) {
        decims[i] = p->st[i].decim;
        tapcounts[i] = p->st[i].ntaps;
    }
    return p->nstages;
}

int orc_decim_taps(int ratio, int stage, float* out, int cap) {
    const decim_plan* p = find_plan(ratio);
    if (!p || stage < 0 || stage >= p->nstages) { return 0; }
    int n = p->st[stage].ntaps;
    if (out) { memcpy(out, p->st[stage].taps, sizeof(float) * (size_t)(n < cap ? n : cap)); }
    return n;
}

/* ---- RationalResampler::reconfigure  (core/src/dsp/multirate/rational_resampler.h:120-165) ---- */
static int igcd(int a, int b) {
    if (a < 0) { a = -a; }
    if (b < 0) { b = -b; }
    while (b) { int t = a % b; a = b; b = t; }
    return a;
}
#define PD_MAX_RATIO (1 << 13) /* PowerDecimator::getMaxRatio, power_decimator.h:28-30 */

typedef struct {
    int mode, useDecim, predecRatio, interp, decim;
    float* rtaps; int nrtaps; /* prototype, already * interp */
} rr_plan;

static void rr_make_plan(double inSR, double outSR, rr_plan* pl) {
    int predecPower = (int)floor(log2(inSR / outSR));
    if (predecPower > PD_MAX_RATIO) { predecPower = PD_MAX_RATIO; }
    int predecRatio = (predecPower >= 0 && predecPower < 31) ? (1 << predecPower) : PD_MAX_RATIO;
    if (predecPower < 0) { predecRatio = 0; } /* 1 << negative is UB in the reference; never reached with useDecim */
    if (predecRatio > PD_MAX_RATIO) { predecRatio = PD_MAX_RATIO; }
    double intSamplerate = inSR;
    int useDecim = (inSR > outSR && predecPower > 0);
    if (useDecim) { intSamplerate = inSR / (double)predecRatio; }
    int IntSR = (int)round(intSamplerate);
    int OutSR = (int)round(outSR);
    int g = igcd(IntSR, OutSR);
    int interp = OutSR / g;
    int decim = IntSR / g;
    pl->useDecim = useDecim;
    pl->predecRatio = useDecim ? predecRatio : 1;
    pl->interp = interp;
    pl->decim = decim;
    pl->rtaps = NULL;
    pl->nrtaps = 0;
    if (interp == decim) {
        pl->mode = useDecim ? 1 : 3;
        return;
    }
    double tapSamplerate = intSamplerate * (double)interp;
    double tapBandwidth = (inSR < outSR ? inSR : outSR) / 2.0;
    double tapTransWidth = tapBandwidth * 0.1;
    pl->rtaps = lowpass_taps(tapBandwidth, tapTransWidth, tapSamplerate, 0, &pl->nrtaps);
    for (int i = 0; i < pl->nrtaps; i++) { pl->rtaps[i] *= (float)interp; }
    pl->mode = useDecim ? 0 : 2;
}

int orc_resamp_plan_get(double inSR, double outSR, orc_resamp_plan* out) {
    rr_plan pl;
    rr_make_plan(inSR, outSR, &pl);
    out->mode = pl.mode;
    out->predec_ratio = pl.predecRatio;
    out->interp = pl.interp;
    out->decim = pl.decim;
    out->ntaps = pl.nrtaps;
    out->taps_per_phase = pl.nrtaps ? (pl.nrtaps + pl.interp - 1) / pl.interp : 0;
    free(pl.rtaps);
    return 0;
}

int orc_resamp_taps(double inSR, double outSR, float* out, int cap) {
    rr_plan pl;
    rr_make_plan(inSR, outSR, &pl);
    int n = pl.nrtaps;
    if (out && n) { memcpy(out, pl.rtaps, sizeof(float) * (size_t)(n < cap ? n : cap)); }
    free(pl.rtaps);
    return n;
}

/* ------------------------------------------------------------------ */
/* streaming blocks                                                    */
/* ------------------------------------------------------------------ */

typedef struct node node;
struct node {
    int (*process)(node*, int, const void*, void*);
    void (*reset)(node*);
    void (*destroy)(node*);
};

/* ---- channel::FrequencyXlator  (core/src/dsp/channel/frequency_xlator.h:15-50) ----
 * Rotator mode 0 (default, the declared oracle): the faithful fp32 recurrence phase *= phaseDelta of VOLK's
 * rotator2.  Mode 1 ("exact phase"): same phaseDelta, but the phase is carried as an fp64 angle
 * n*angle(phaseDelta) and rounded to fp32 per sample, i.e. the recurrence WITHOUT its fp32 rounding walk.
 * SSB/DSB audio Re{x e^{j theta}} is first-order sensitive to that walk, so those outputs
 * are gated against mode 1 and the mode-0-vs-mode-1 gap is reported as the reference's own numerical floor.
 * Mode 1 exists only in this restatement (the reference-header build is always mode 0). */
static int g_rotator_mode = 0;
void orc_set_rotator_mode(int mode) { g_rotator_mode = mode; }
typedef struct { cf32 phase, delta; double angle; } xlator_t;
static void xl_init(xlator_t* x, double offsetRad) {
    x->phase.re = 1.0f; x->phase.im = 0.0f;
    x->delta.re = (float)cos(offsetRad);
    x->delta.im = (float)sin(offsetRad);
    x->angle = 0.0;
}
static void xl_set(xlator_t* x, double offsetRad) {
    x->delta.re = (float)cos(offsetRad);
    x->delta.im = (float)sin(offsetRad);
}
static int xl_process(xlator_t* x, int count, const cf32* in, cf32* out) {
    if (g_rotator_mode == 0) {
        ovk_rotator2(out, in, &x->delta, &x->phase, (unsigned)count);
        return count;
    }
    const double w = atan2((double)x->delta.im, (double)x->delta.re);
    for (int i = 0; i < count; i++) {
        double a = x->angle + w * (double)i;
        float pr = (float)cos(a), pi = (float)sin(a);
        cf32 v = in[i], y;
        y.re = v.re * pr - v.im * pi;
        y.im = v.re * pi + v.im * pr;
        out[i] = y;
    }
    x->angle = fmod(x->angle + w * (double)count, 2.0 * DB_M_PI);
    return count;
}

/* ---- filter::FIR / DecimatingFIR delay line  (fir.h:20-29,62-83; decimating_fir.h:45-68) ---- */
typedef struct {
    int ntaps, decim, offset, esize; /* esize: floats per element (1 real, 2 complex) */
    float* taps;
    float* buffer; /* (STREAM_BUFFER_SIZE + 64000) elements */
} fir_t;

static void fir_init(fir_t* f, const float* taps, int n, int decim, int esize) {
    f->ntaps = n;
    f->decim = decim;
    f->offset = 0;
    f->esize = esize;
    f->taps = (float*)malloc(sizeof(float) * (size_t)n);
    memcpy(f->taps, taps, sizeof(float) * (size_t)n);
    f->buffer = (float*)calloc((size_t)(STREAM_BUFFER_SIZE + DELAY_EXTRA) * (size_t)esize, sizeof(float));
}
static void fir_free(fir_t* f) { free(f->taps); free(f->buffer); }
static void fir_reset(fir_t* f) {
    memset(f->buffer, 0, sizeof(float) * (size_t)f->esize * (size_t)(f->ntaps - 1));
    f->offset = 0;
}
/* FIR::setTaps  (fir.h:31-52): keeps the most recent history */
static void fir_set_taps(fir_t* f, const float* taps, int n) {
    int oldTC = f->ntaps;
    free(f->taps);
    f->taps = (float*)malloc(sizeof(float) * (size_t)n);
    memcpy(f->taps, taps, sizeof(float) * (size_t)n);
    f->ntaps = n;
    size_t es = sizeof(float) * (size_t)f->esize;
    if (n < oldTC) {
        memmove(f->buffer, f->buffer + (size_t)(oldTC - n) * f->esize, (size_t)(n - 1) * es);
    }
    else if (n > oldTC) {
        memmove(f->buffer + (size_t)(n - oldTC) * f->esize, f->buffer, (size_t)(oldTC - 1) * es);
        memset(f->buffer, 0, (size_t)(n - oldTC) * es);
    }
}

/* complex data, real taps: FIR<complex_t,float> with optional decimation */
static int fir_process_c(fir_t* f, int count, const cf32* in, cf32* out) {
    cf32* buf = (cf32*)f->buffer;
    memcpy(&buf[f->ntaps - 1], in, sizeof(cf32) * (size_t)count);
    int outCount = 0;
    if (f->decim == 1) {
        for (int i = 0; i < count; i++) { ovk_dot_32fc_32f(&out[i], &buf[i], f->taps, (unsigned)f->ntaps); }
        outCount = count;
    }
    else {
        for (; f->offset < count; f->offset += f->decim) {
            ovk_dot_32fc_32f(&out[outCount++], &buf[f->offset], f->taps, (unsigned)f->ntaps);
        }
        f->offset -= count;
    }
    memmove(buf, &buf[count], sizeof(cf32) * (size_t)(f->ntaps - 1));
    return outCount;
}
/* real data, real taps: FIR<float,float> */
static int fir_process_r(fir_t* f, int count, const float* in, float* out) {
    float* buf = f->buffer;
    memcpy(&buf[f->ntaps - 1], in, sizeof(float) * (size_t)count);
    for (int i = 0; i < count; i++) { ovk_dot_32f(&out[i], &buf[i], f->taps, (unsigned)f->ntaps); }
    memmove(buf, &buf[count], sizeof(float) * (size_t)(f->ntaps - 1));
    return count;
}

/* ---- multirate::PowerDecimator  (core/src/dsp/multirate/power_decimator.h:51-67,92-110) ---- */
typedef struct { int ratio, nstages; fir_t st[MAX_STAGES]; } pdecim_t;
static int pd_init(pdecim_t* d, int ratio) {
    d->ratio = ratio;
    d->nstages = 0;
    if (ratio == 1) { return 0; }
    const decim_plan* p = find_plan(ratio);
    if (!p) { return -1; }
    d->nstages = p->nstages;
    for (int i = 0; i < p->nstages; i++) { fir_init(&d->st[i], p->st[i].taps, p->st[i].ntaps, p->st[i].decim, 2); }
    return 0;
}
static void pd_free(pdecim_t* d) { for (int i = 0; i < d->nstages; i++) { fir_free(&d->st[i]); } d->nstages = 0; }
static void pd_reset(pdecim_t* d) { for (int i = 0; i < d->nstages; i++) { fir_reset(&d->st[i]); } }
static int pd_process(pdecim_t* d, int count, const cf32* in, cf32* out) {
    if (d->ratio == 1) {
        memcpy(out, in, sizeof(cf32) * (size_t)count);
        return count;
    }
    const cf32* data = in;
    for (int i = 0; i < d->nstages; i++) {
        count = fir_process_c(&d->st[i], count, data, out);
        data = out;
    }
    return count;
}

/* ---- multirate::PolyphaseResampler + buildPolyphaseBank
 *      (polyphase_resampler.h:20-34,69-99; polyphase_bank.h:15-48) ---- */
typedef struct {
    int interp, decim, tapsPerPhase, phase, offset;
    float** phases;
    cf32* buffer;
} ppr_t;
static void ppr_build(ppr_t* r, int interp, int decim, const float* taps, int ntaps) {
    r->interp = interp;
    r->decim = decim;
    r->tapsPerPhase = (ntaps + interp - 1) / interp;
    r->phases = (float**)malloc(sizeof(float*) * (size_t)interp);
    for (int i = 0; i < interp; i++) { r->phases[i] = (float*)calloc((size_t)r->tapsPerPhase, sizeof(float)); }
    int tot = interp * r->tapsPerPhase;
    for (int i = 0; i < tot; i++) {
        r->phases[(interp - 1) - (i % interp)][i / interp] = (i < ntaps) ? taps[i] : 0;
    }
}
static void ppr_init(ppr_t* r, int interp, int decim, const float* taps, int ntaps) {
    ppr_build(r, interp, decim, taps, ntaps);
    r->buffer = (cf32*)calloc((size_t)(STREAM_BUFFER_SIZE + DELAY_EXTRA), sizeof(cf32));
    r->phase = 0;
    r->offset = 0;
}
static void ppr_free(ppr_t* r) {
    for (int i = 0; i < r->interp; i++) { free(r->phases[i]); }
    free(r->phases);
    free(r->buffer);
}
static void ppr_reset(ppr_t* r) {
    memset(r->buffer, 0, sizeof(cf32) * (size_t)(r->tapsPerPhase - 1));
    r->phase = 0;
    r->offset = 0;
}
static int ppr_process(ppr_t* r, int count, const cf32* in, cf32* out) {
    int outCount = 0;
    memcpy(&r->buffer[r->tapsPerPhase - 1], in, sizeof(cf32) * (size_t)count);
    while (r->offset < count) {
        ovk_dot_32fc_32f(&out[outCount++], &r->buffer[r->offset], r->phases[r->phase], (unsigned)r->tapsPerPhase);
        r->phase += r->decim;
        r->offset += r->phase / r->interp;
        r->phase = r->phase % r->interp;
    }
    r->offset -= count;
    memmove(r->buffer, &r->buffer[count], sizeof(cf32) * (size_t)(r->tapsPerPhase - 1));
    return outCount;
}

/* ---- multirate::RationalResampler  (rational_resampler.h:82-96,120-165).  complex_t and stereo_t use the
 *      same complex-data x real-taps dot product (polyphase_resampler.h:80-82), so one implementation. ---- */
typedef struct { int mode; int hasDecim, hasResamp; pdecim_t decim; ppr_t resamp; } rresamp_t;
static int rr_init(rresamp_t* r, double inSR, double outSR) {
    rr_plan pl;
    rr_make_plan(inSR, outSR, &pl);
    r->mode = pl.mode;
    r->hasDecim = r->hasResamp = 0;
    if (pl.useDecim) {
        if (pd_init(&r->decim, pl.predecRatio)) { free(pl.rtaps); return -1; }
        r->hasDecim = 1;
    }
    if (pl.nrtaps) {
        ppr_init(&r->resamp, pl.interp, pl.decim, pl.rtaps, pl.nrtaps);
        r->hasResamp = 1;
    }
    free(pl.rtaps);
    return 0;
}
static void rr_free(rresamp_t* r) {
    if (r->hasDecim) { pd_free(&r->decim); }
    if (r->hasResamp) { ppr_free(&r->resamp); }
}
static void rr_reset(rresamp_t* r) {
    if (r->hasDecim) { pd_reset(&r->decim); }
    if (r->hasResamp) { ppr_reset(&r->resamp); }
}
static int rr_process(rresamp_t* r, int count, const cf32* in, cf32* out) {
    switch (r->mode) {
    case 0:
        count = pd_process(&r->decim, count, in, out);
        return ppr_process(&r->resamp, count, out, out);
    case 1:
        return pd_process(&r->decim, count, in, out);
    case 2:
        return ppr_process(&r->resamp, count, in, out);
    default:
        memcpy(out, in, sizeof(cf32) * (size_t)count);
        return count;
    }
}

/* ---- channel::RxVFO  (core/src/dsp/channel/rx_vfo.h:17-31,60-77,89-100,117-121) ---- */
typedef struct {
    double inSR, outSR, bw, offset;
    int filterNeeded;
    xlator_t xl;
    rresamp_t rs;
    fir_t filt;
} rxvfo_t;
static float* vfo_taps(double bw, double outSR, int* n) {
    double filterWidth = bw / 2.0;
    return lowpass_taps(filterWidth, filterWidth * 0.1, outSR, 0, n);
}
static int vfo_process(rxvfo_t* v, int count, const cf32* in, cf32* out) {
    xl_process(&v->xl, count, in, out);
    if (!v->filterNeeded) { return rr_process(&v->rs, count, out, out); }
    count = rr_process(&v->rs, count, out, out);
    fir_process_c(&v->filt, count, out, out);
    return count;
}

/* ---- demod::Quadrature  (core/src/dsp/demod/quadrature.h:19-46), math::normalizePhase
 *      (math/normalize_phase.h:6-10) ---- */
typedef struct { float invDeviation, phase; } fmquad_t;
static void quad_init(fmquad_t* q, double deviationHz, double sr) {
    q->invDeviation = (float)(1.0 / hz_to_rads(deviationHz, sr));
    q->phase = 0.0f;
}
static int quad_process(fmquad_t* q, int count, const cf32* in, float* out) {
    for (int i = 0; i < count; i++) {
        float cphase = atan2f(in[i].im, in[i].re);
        float diff = cphase - q->phase;
        if (diff > FL_M_PI) { diff -= 2.0f * FL_M_PI; }
        else if (diff <= -FL_M_PI) { diff += 2.0f * FL_M_PI; }
        out[i] = diff * q->invDeviation;
        q->phase = cphase;
    }
    return count;
}

/* ---- loop::AGC<T>  (core/src/dsp/loop/agc.h:13-24,70-110) ---- */
typedef struct { float setPoint, attack, invAttack, decay, invDecay, maxGain, maxOutputAmp, initGain, amp; } agc_t;
static void agc_init(agc_t* a, double setPoint, double attack, double decay, double maxGain, double maxOut, double initGain) {
    a->setPoint = (float)setPoint;
    a->attack = (float)attack;
    a->invAttack = 1.0f - a->attack;
    a->decay = (float)decay;
    a->invDecay = 1.0f - a->decay;
    a->maxGain = (float)maxGain;
    a->maxOutputAmp = (float)maxOut;
    a->initGain = (float)initGain;
    a->amp = a->setPoint / a->initGain;
}
static void agc_reset(agc_t* a) { a->amp = a->setPoint / a->initGain; }
static float agc_gain_step(agc_t* a, float inAmp) {
    float gain;
    if (inAmp != 0.0f) {
        a->amp = (inAmp > a->amp) ? ((a->amp * a->invAttack) + (inAmp * a->attack)) : ((a->amp * a->invDecay) + (inAmp * a->decay));
        gain = a->setPoint / a->amp;
        if (a->maxGain < gain) { gain = a->maxGain; }
    }
    else {
        gain = 1.0f;
    }
    return gain;
}
static int agc_process_f(agc_t* a, int count, const float* in, float* out) {
    for (int i = 0; i < count; i++) {
        float inAmp = fabsf(in[i]);
        float gain = agc_gain_step(a, inAmp);
        if (inAmp * gain > a->maxOutputAmp) {
            float maxAmp = 0;
            for (int j = i; j < count; j++) {
                inAmp = fabsf(in[j]);
                if (inAmp > maxAmp) { maxAmp = inAmp; }
            }
            a->amp = maxAmp;
            gain = a->setPoint / a->amp;
            if (a->maxGain < gain) { gain = a->maxGain; }
        }
        out[i] = in[i] * gain;
    }
    return count;
}
static float camp(cf32 x) { return sqrtf((x.re * x.re) + (x.im * x.im)); } /* complex_t::amplitude, types.h:86-88 */
static int agc_process_c(agc_t* a, int count, const cf32* in, cf32* out) {
    for (int i = 0; i < count; i++) {
        float inAmp = camp(in[i]);
        float gain = agc_gain_step(a, inAmp);
        if (inAmp * gain > a->maxOutputAmp) {
            float maxAmp = 0;
            for (int j = i; j < count; j++) {
                inAmp = camp(in[j]);
                if (inAmp > maxAmp) { maxAmp = inAmp; }
            }
            a->amp = maxAmp;
            gain = a->setPoint / a->amp;
            if (a->maxGain < gain) { gain = a->maxGain; }
        }
        cf32 x = in[i];
        out[i].re = x.re * gain;
        out[i].im = x.im * gain;
    }
    return count;
}

/* ---- correction::DCBlocker<T>  (core/src/dsp/correction/dc_blocker.h:15-24,54-60) ---- */
static int dcblock_f(float rate, float* offset, int count, const float* in, float* out) {
    float o = *offset;
    for (int i = 0; i < count; i++) {
        out[i] = in[i] - o;
        o += out[i] * rate;
    }
    *offset = o;
    return count;
}
static int dcblock_c(float rate, cf32* offset, int count, const cf32* in, cf32* out) {
    cf32 o = *offset;
    for (int i = 0; i < count; i++) {
        cf32 y = { in[i].re - o.re, in[i].im - o.im };
        out[i] = y;
        o.re += y.re * rate;
        o.im += y.im * rate;
    }
    *offset = o;
    return count;
}

/* ------------------------------------------------------------------ */
/* node wrappers                                                       */
/* ------------------------------------------------------------------ */
#define NODE_ALLOC(T) T* n = (T*)calloc(1, sizeof(T))

typedef struct { node base; xlator_t x; } n_xl;
static int n_xl_proc(node* b, int c, const void* i, void* o) { return xl_process(&((n_xl*)b)->x, c, (const cf32*)i, (cf32*)o); }
static void n_xl_reset(node* b) { ((n_xl*)b)->x.phase.re = 1.0f; ((n_xl*)b)->x.phase.im = 0.0f; ((n_xl*)b)->x.angle = 0.0; }
static void n_plain_destroy(node* b) { free(b); }
void* orc_xlator_create(double offsetHz, double sr) {
    NODE_ALLOC(n_xl);
    n->base.process = n_xl_proc; n->base.reset = n_xl_reset; n->base.destroy = n_plain_destroy;
    xl_init(&n->x, hz_to_rads(offsetHz, sr));
    return n;
}
void orc_xlator_set_offset(void* h, double offsetHz, double sr) { xl_set(&((n_xl*)h)->x, hz_to_rads(offsetHz, sr)); }
void orc_xlator_get_phase(void* h, float* ph, float* dl) {
    n_xl* n = (n_xl*)h;
    ph[0] = n->x.phase.re; ph[1] = n->x.phase.im;
    dl[0] = n->x.delta.re; dl[1] = n->x.delta.im;
}

typedef struct { node base; pdecim_t d; } n_pd;
static int n_pd_proc(node* b, int c, const void* i, void* o) { return pd_process(&((n_pd*)b)->d, c, (const cf32*)i, (cf32*)o); }
static void n_pd_reset(node* b) { pd_reset(&((n_pd*)b)->d); }
static void n_pd_destroy(node* b) { pd_free(&((n_pd*)b)->d); free(b); }
void* orc_decim_create(int ratio) {
    NODE_ALLOC(n_pd);
    n->base.process = n_pd_proc; n->base.reset = n_pd_reset; n->base.destroy = n_pd_destroy;
    if (pd_init(&n->d, ratio)) { free(n); return NULL; }
    return n;
}

typedef struct { node base; rresamp_t r; } n_rr;
static int n_rr_proc(node* b, int c, const void* i, void* o) { return rr_process(&((n_rr*)b)->r, c, (const cf32*)i, (cf32*)o); }
static void n_rr_reset(node* b) { rr_reset(&((n_rr*)b)->r); }
static void n_rr_destroy(node* b) { rr_free(&((n_rr*)b)->r); free(b); }
void* orc_resamp_create(double inSR, double outSR) {
    NODE_ALLOC(n_rr);
    n->base.process = n_rr_proc; n->base.reset = n_rr_reset; n->base.destroy = n_rr_destroy;
    if (rr_init(&n->r, inSR, outSR)) { free(n); return NULL; }
    return n;
}
void* orc_resamp_stereo_create(double inSR, double outSR) { return orc_resamp_create(inSR, outSR); }

typedef struct { node base; fir_t f; } n_fir;
static int n_firc_proc(node* b, int c, const void* i, void* o) { return fir_process_c(&((n_fir*)b)->f, c, (const cf32*)i, (cf32*)o); }
static int n_firr_proc(node* b, int c, const void* i, void* o) { return fir_process_r(&((n_fir*)b)->f, c, (const float*)i, (float*)o); }
static void n_fir_reset(node* b) { fir_reset(&((n_fir*)b)->f); }
static void n_fir_destroy(node* b) { fir_free(&((n_fir*)b)->f); free(b); }
void* orc_fir_cr_create(const float* taps, int nt) {
    NODE_ALLOC(n_fir);
    n->base.process = n_firc_proc; n->base.reset = n_fir_reset; n->base.destroy = n_fir_destroy;
    fir_init(&n->f, taps, nt, 1, 2);
    return n;
}
void* orc_fir_rr_create(const float* taps, int nt) {
    NODE_ALLOC(n_fir);
    n->base.process = n_firr_proc; n->base.reset = n_fir_reset; n->base.destroy = n_fir_destroy;
    fir_init(&n->f, taps, nt, 1, 1);
    return n;
}
void* orc_decfir_cr_create(const float* taps, int nt, int decim) {
    NODE_ALLOC(n_fir);
    n->base.process = n_firc_proc; n->base.reset = n_fir_reset; n->base.destroy = n_fir_destroy;
    fir_init(&n->f, taps, nt, decim, 2);
    /* a DecimatingFIR with decimation 1 still walks `offset` (decimating_fir.h:50-62): identical result */
    return n;
}

typedef struct { node base; rxvfo_t v; } n_vfo;
static int n_vfo_proc(node* b, int c, const void* i, void* o) { return vfo_process(&((n_vfo*)b)->v, c, (const cf32*)i, (cf32*)o); }
static void n_vfo_reset(node* b) {
    rxvfo_t* v = &((n_vfo*)b)->v;
    v->xl.phase.re = 1.0f; v->xl.phase.im = 0.0f; v->xl.angle = 0.0;
    rr_reset(&v->rs);
    fir_reset(&v->filt);
}
static void n_vfo_destroy(node* b) { rxvfo_t* v = &((n_vfo*)b)->v; rr_free(&v->rs); fir_free(&v->filt); free(b); }
void* orc_rxvfo_create(double inSR, double outSR, double bw, double offset) {
    NODE_ALLOC(n_vfo);
    n->base.process = n_vfo_proc; n->base.reset = n_vfo_reset; n->base.destroy = n_vfo_destroy;
    rxvfo_t* v = &n->v;
    v->inSR = inSR; v->outSR = outSR; v->bw = bw; v->offset = offset;
    v->filterNeeded = (bw != outSR);
    xl_init(&v->xl, hz_to_rads(-offset, inSR));
    if (rr_init(&v->rs, inSR, outSR)) { free(n); return NULL; }
    int nt;
    float* t = vfo_taps(bw, outSR, &nt);
    fir_init(&v->filt, t, nt, 1, 2);
    free(t);
    return n;
}
void orc_rxvfo_set_offset(void* h, double offset) {
    rxvfo_t* v = &((n_vfo*)h)->v;
    v->offset = offset;
    xl_set(&v->xl, hz_to_rads(-offset, v->inSR));
}
void orc_rxvfo_set_bandwidth(void* h, double bw) {
    rxvfo_t* v = &((n_vfo*)h)->v;
    v->bw = bw;
    v->filterNeeded = (bw != v->outSR);
    if (v->filterNeeded) {
        int nt;
        float* t = vfo_taps(bw, v->outSR, &nt);
        fir_set_taps(&v->filt, t, nt);
        free(t);
    }
}

typedef struct { node base; fmquad_t q; } n_quad;
static int n_quad_proc(node* b, int c, const void* i, void* o) { return quad_process(&((n_quad*)b)->q, c, (const cf32*)i, (float*)o); }
static void n_quad_reset(node* b) { ((n_quad*)b)->q.phase = 0.0f; }
void* orc_quad_create(double dev, double sr) {
    NODE_ALLOC(n_quad);
    n->base.process = n_quad_proc; n->base.reset = n_quad_reset; n->base.destroy = n_plain_destroy;
    quad_init(&n->q, dev, sr);
    return n;
}

/* scratch for mono audio */
typedef struct { float* p; int cap; } scratch_t;
static float* scratch_get(scratch_t* s, int count) {
    if (s->cap < count) { free(s->p); s->p = (float*)malloc(sizeof(float) * (size_t)count); s->cap = count; }
    return s->p;
}
/* convert::MonoToStereo / LRToStereo(m,m)  (convert/mono_to_stereo.h:13, l_r_to_stereo.h:21) */
static void mono_to_stereo(int count, const float* m, float* out) {
    for (int i = 0; i < count; i++) { out[2 * i] = m[i]; out[2 * i + 1] = m[i]; }
}

/* ---- demod::BroadcastFM mono branch  (core/src/dsp/demod/broadcast_fm.h:36-52,144-147,192-212) ---- */
typedef struct { node base; fmquad_t q; fir_t al; int lowPass; scratch_t s; } n_wfm;
static int n_wfm_proc(node* b, int count, const void* in, void* out) {
    n_wfm* w = (n_wfm*)b;
    float* m = scratch_get(&w->s, count);
    quad_process(&w->q, count, (const cf32*)in, m);
    if (w->lowPass) { fir_process_r(&w->al, count, m, m); }
    mono_to_stereo(count, m, (float*)out);
    return count;
}
static void n_wfm_reset(node* b) { n_wfm* w = (n_wfm*)b; w->q.phase = 0.0f; fir_reset(&w->al); }
static void n_wfm_destroy(node* b) { n_wfm* w = (n_wfm*)b; fir_free(&w->al); free(w->s.p); free(b); }
/* ---- demod::BroadcastFM, RDS side output  (core/src/dsp/demod/broadcast_fm.h:52-53,165-170,196-202): discriminator ->
 *      RealToComplex -> FrequencyXlator(-57 kHz) -> RationalResampler to 5 kS/s; the same three blocks in the mono and in
 *      the stereo branch ---- */
typedef struct { node base; fmquad_t q; xlator_t x; rresamp_t r; scratch_t s; cf32* c; size_t cap; } n_wfmrds;
static int n_wfmrds_proc(node* b, int count, const void* in, void* out) {
    n_wfmrds* w = (n_wfmrds*)b;
    float* m = scratch_get(&w->s, count);
    if ((size_t)count > w->cap) { w->cap = (size_t)count + 1024; w->c = (cf32*)realloc(w->c, w->cap * sizeof(cf32)); }
    quad_process(&w->q, count, (const cf32*)in, m);
    for (int i = 0; i < count; i++) { w->c[i].re = m[i]; w->c[i].im = 0.0f; }       /* convert::RealToComplex (real_to_complex.h) */
    xl_process(&w->x, count, w->c, w->c);
    return rr_process(&w->r, count, w->c, (cf32*)out);
}
static void n_wfmrds_reset(node* b) { (void)b; }
static void n_wfmrds_destroy(node* b) { n_wfmrds* w = (n_wfmrds*)b; rr_free(&w->r); free(w->s.p); free(w->c); free(b); }

/* ---- demod::BroadcastFM stereo branch  (core/src/dsp/demod/broadcast_fm.h:36-66,147-190): RealToComplex -> pilot band-pass
 *      (FIR<complex_t,complex_t>, taps::bandPass<complex_t>(18750,19250,3000,fs,odd)) -> loop::PLL (pll.h:64-70 over
 *      PhaseControlLoop, phase_control_loop.h:27-32,58-85) -> Delay x2 -> conj, two complex multiplies -> real part * 2 ->
 *      L = (L+R) + (L-R), R = (L+R) - (L-R) -> two audio low-passes -> LRToStereo.  RDS output not restated (off). ---- */
/* math::normalizePhase  (math/normalize_phase.h:6-10) */
static float normalize_phase(float diff) {
    if (diff > FL_M_PI) { diff -= 2.0f * FL_M_PI; }
    else if (diff <= -FL_M_PI) { diff += 2.0f * FL_M_PI; }
    return diff;
}
typedef struct { float alpha, beta, phase, freq, minFreq, maxFreq; } pll_t;
static void pll_init(pll_t* p, double bandwidth, double initPhase, double initFreq, double minFreq, double maxFreq) {
    /* PhaseControlLoop<float>::criticallyDamped with T = float: the arguments are narrowed to float first */
    float bw = (float)bandwidth;
    float damp = (float)(sqrt(2.0) / 2.0);
    float den = (float)(1.0 + 2.0 * damp * bw + bw * bw);
    p->alpha = (4 * damp * bw) / den;
    p->beta = (4 * bw * bw) / den;
    p->phase = (float)initPhase;
    p->freq = (float)initFreq;
    p->minFreq = (float)minFreq;
    p->maxFreq = (float)maxFreq;
}
static void pll_process(pll_t* p, int count, const cf32* in, cf32* out) {
    const float minPhase = -FL_M_PI, maxPhase = FL_M_PI, phaseDelta = maxPhase - minPhase;
    for (int i = 0; i < count; i++) {
        out[i].re = cosf(p->phase);                                  /* math::phasor */
        out[i].im = sinf(p->phase);
        float err = normalize_phase(atan2f(in[i].im, in[i].re) - p->phase);
        p->freq += p->beta * err;                                    /* PhaseControlLoop::advance */
        if (p->freq > p->maxFreq) { p->freq = p->maxFreq; }
        else if (p->freq < p->minFreq) { p->freq = p->minFreq; }
        p->phase += p->freq + (p->alpha * err);
        while (p->phase > maxPhase) { p->phase -= phaseDelta; }
        while (p->phase < minPhase) { p->phase += phaseDelta; }
    }
}
/* math::Delay<T>  (math/delay.h:43-53) on `es` floats per element */
typedef struct { int delay, es; float* buf; } delay_t;
static void delay_init(delay_t* d, int delay, int es) {
    d->delay = delay; d->es = es;
    d->buf = (float*)calloc((size_t)(STREAM_BUFFER_SIZE + DELAY_EXTRA) * (size_t)es, sizeof(float));
}
static void delay_process(delay_t* d, int count, const float* in, float* out) {
    size_t e = sizeof(float) * (size_t)d->es;
    memcpy(d->buf + (size_t)d->delay * d->es, in, e * (size_t)count);
    memcpy(out, d->buf, e * (size_t)count);
    memmove(d->buf, d->buf + (size_t)count * d->es, e * (size_t)d->delay);
}
/* complex data, complex taps: FIR<complex_t,complex_t>  (fir.h:74-76: volk_32fc_x2_dot_prod_32fc) */
static int fir_process_cc(fir_t* f, int count, const cf32* in, cf32* out) {
    cf32* buf = (cf32*)f->buffer;
    memcpy(&buf[f->ntaps - 1], in, sizeof(cf32) * (size_t)count);
    for (int i = 0; i < count; i++) { ovk_dot_32fc_32fc((ovk_cf32*)&out[i], (const ovk_cf32*)&buf[i], (const ovk_cf32*)f->taps, (unsigned)f->ntaps); }
    memmove(buf, &buf[count], sizeof(cf32) * (size_t)(f->ntaps - 1));
    return count;
}
typedef struct {
    node base; fmquad_t q; fir_t pilot, al, ar; pll_t pll; delay_t lpr, lmr; int lowPass;
    scratch_t sm, sc, sp, sv, sd, sl, sr;
    double sr_hz;
} n_wfms;
static int n_wfms_proc(node* b, int count, const void* in, void* out) {
    n_wfms* w = (n_wfms*)b;
    float* m = scratch_get(&w->sm, count);
    cf32* z = (cf32*)scratch_get(&w->sc, 2 * count);
    cf32* pf = (cf32*)scratch_get(&w->sp, 2 * count);
    cf32* vco = (cf32*)scratch_get(&w->sv, 2 * count);
    cf32* zd = (cf32*)scratch_get(&w->sd, 2 * count);
    float* l = scratch_get(&w->sl, count);
    float* r = scratch_get(&w->sr, count);
    quad_process(&w->q, count, (const cf32*)in, m);
    for (int i = 0; i < count; i++) { z[i].re = m[i]; z[i].im = 0.0f; }          /* RealToComplex (interleave with zeros) */
    fir_process_cc(&w->pilot, count, z, pf);
    pll_process(&w->pll, count, pf, vco);
    delay_process(&w->lpr, count, m, m);
    delay_process(&w->lmr, count, (const float*)z, (float*)zd);
    for (int i = 0; i < count; i++) { vco[i].im = -vco[i].im; }                    /* math::Conjugate */
    ovk_mul_32fc_32fc((ovk_cf32*)zd, (const ovk_cf32*)zd, (const ovk_cf32*)vco, (unsigned)count);
    ovk_mul_32fc_32fc((ovk_cf32*)zd, (const ovk_cf32*)zd, (const ovk_cf32*)vco, (unsigned)count);
    for (int i = 0; i < count; i++) {
        float lmr = zd[i].re * 2.0f;                                             /* ComplexToReal, x2 */
        l[i] = m[i] + lmr;
        r[i] = m[i] - lmr;
    }
    if (w->lowPass) {
        fir_process_r(&w->al, count, l, l);
        fir_process_r(&w->ar, count, r, r);
    }
    float* o = (float*)out;
    for (int i = 0; i < count; i++) { o[2 * i] = l[i]; o[2 * i + 1] = r[i]; }     /* LRToStereo */
    return count;
}
static void n_wfms_reset(node* b) { (void)b; }   /* BroadcastFM::reset only resets demod / FIRs (not used by the tests) */
static void n_wfms_destroy(node* b) {
    n_wfms* w = (n_wfms*)b;
    fir_free(&w->pilot); fir_free(&w->al); fir_free(&w->ar); free(w->lpr.buf); free(w->lmr.buf);
    free(w->sm.p); free(w->sc.p); free(w->sp.p); free(w->sv.p); free(w->sd.p); free(w->sl.p); free(w->sr.p);
    free(b);
}
static void* wfm_stereo_create(double dev, double sr, int lowPass) {
    NODE_ALLOC(n_wfms);
    n->base.process = n_wfms_proc; n->base.reset = n_wfms_reset; n->base.destroy = n_wfms_destroy;
    quad_init(&n->q, dev, sr);
    int np = orc_bandpass_c(18750.0, 19250.0, 3000.0, sr, 1, NULL, 0);
    float* pt = (float*)malloc(sizeof(float) * 2 * (size_t)np);
    orc_bandpass_c(18750.0, 19250.0, 3000.0, sr, 1, pt, np);
    fir_init(&n->pilot, pt, 2 * np, 1, 2);          /* taps stored as 2*np floats; ntaps fixed up below */
    n->pilot.ntaps = np;
    free(pt);
    pll_init(&n->pll, 25000.0 / sr, 0.0, hz_to_rads(19000.0, sr), hz_to_rads(18750.0, sr), hz_to_rads(19250.0, sr));
    delay_init(&n->lpr, ((np - 1) / 2) + 1, 1);
    delay_init(&n->lmr, ((np - 1) / 2) + 1, 2);
    int nt;
    float* t = lowpass_taps(15000.0, 4000.0, sr, 0, &nt);
    fir_init(&n->al, t, nt, 1, 1);
    fir_init(&n->ar, t, nt, 1, 1);
    free(t);
    n->lowPass = lowPass;
    return n;
}
void* orc_wfm_rds_create(double dev, double sr) {
    NODE_ALLOC(n_wfmrds);
    n->base.process = n_wfmrds_proc; n->base.reset = n_wfmrds_reset; n->base.destroy = n_wfmrds_destroy;
    quad_init(&n->q, dev, sr);
    xl_init(&n->x, hz_to_rads(-57000.0, sr));
    if (rr_init(&n->r, sr, 5000.0)) { free(n); return NULL; }
    return n;
}
void* orc_wfm_create(double dev, double sr, int stereo, int lowPass) {
    if (stereo) { return wfm_stereo_create(dev, sr, lowPass); }
    NODE_ALLOC(n_wfm);
    n->base.process = n_wfm_proc; n->base.reset = n_wfm_reset; n->base.destroy = n_wfm_destroy;
    quad_init(&n->q, dev, sr);
    int nt;
    float* t = lowpass_taps(15000.0, 4000.0, sr, 0, &nt);
    fir_init(&n->al, t, nt, 1, 1);
    free(t);
    n->lowPass = lowPass;
    return n;
}

/* ---- demod::FM<stereo_t>  (core/src/dsp/demod/fm.h:24-40,79-96,109-134) ---- */
void* orc_nfm_create(double sr, double bw, int lowPass) {
    NODE_ALLOC(n_wfm);
    n->base.process = n_wfm_proc; n->base.reset = n_wfm_reset; n->base.destroy = n_wfm_destroy;
    quad_init(&n->q, bw / 2.0, sr);
    if (lowPass) {
        int nt;
        float* t = lowpass_taps(bw / 2.0, (bw / 2.0) * 0.1, sr, 0, &nt);
        fir_init(&n->al, t, nt, 1, 1);
        free(t);
    }
    else {
        float one = 1.0f; /* loadDummyTaps, fm.h:136-139: the FIR still runs with a single unit tap */
        fir_init(&n->al, &one, 1, 1, 1);
    }
    n->lowPass = 1;
    return n;
}

/* ---- demod::AM<stereo_t>  (core/src/dsp/demod/am.h:28-45,101-133) ---- */
typedef struct {
    node base; int agcMode; agc_t carrier, audio; float dcRate, dcOffset; fir_t lpf; scratch_t s; scratch_t sc;
} n_am;
static int n_am_proc(node* b, int count, const void* in, void* out) {
    n_am* a = (n_am*)b;
    const cf32* x = (const cf32*)in;
    if (a->agcMode == 0) {
        cf32* c = (cf32*)scratch_get(&a->sc, 2 * count);
        agc_process_c(&a->carrier, count, x, c);
        x = c;
    }
    float* m = scratch_get(&a->s, count);
    ovk_magnitude(m, x, (unsigned)count);
    dcblock_f(a->dcRate, &a->dcOffset, count, m, m);
    if (a->agcMode == 1) { agc_process_f(&a->audio, count, m, m); }
    fir_process_r(&a->lpf, count, m, m);
    mono_to_stereo(count, m, (float*)out);
    return count;
}
static void n_am_reset(node* b) { n_am* a = (n_am*)b; agc_reset(&a->carrier); agc_reset(&a->audio); a->dcOffset = 0.0f; }
static void n_am_destroy(node* b) { n_am* a = (n_am*)b; fir_free(&a->lpf); free(a->s.p); free(a->sc.p); free(b); }
void* orc_am_create(int agcMode, double bw, double attack, double decay, double dcRate, double sr) {
    if (agcMode != 0 && agcMode != 1) { return NULL; }
    NODE_ALLOC(n_am);
    n->base.process = n_am_proc; n->base.reset = n_am_reset; n->base.destroy = n_am_destroy;
    n->agcMode = agcMode;
    agc_init(&n->carrier, 1.0, attack, decay, 10e6, 10.0, INFINITY);
    agc_init(&n->audio, 1.0, attack, decay, 10e6, 10.0, INFINITY);
    n->dcRate = (float)dcRate;
    n->dcOffset = 0.0f;
    int nt;
    float* t = lowpass_taps(bw / 2.0, (bw / 2.0) * 0.1, sr, 0, &nt);
    fir_init(&n->lpf, t, nt, 1, 1);
    free(t);
    return n;
}

/* ---- demod::SSB<stereo_t>  (core/src/dsp/demod/ssb.h:22-35,77-92,106-116) ---- */
typedef struct { node base; xlator_t x; agc_t agc; scratch_t s; scratch_t sc; } n_ssb;
static int n_ssb_proc(node* b, int count, const void* in, void* out) {
    n_ssb* s = (n_ssb*)b;
    cf32* c = (cf32*)scratch_get(&s->sc, 2 * count);
    xl_process(&s->x, count, (const cf32*)in, c);
    float* m = scratch_get(&s->s, count);
    for (int i = 0; i < count; i++) { m[i] = c[i].re; } /* convert::ComplexToReal */
    agc_process_f(&s->agc, count, m, m);
    mono_to_stereo(count, m, (float*)out);
    return count;
}
static void n_ssb_reset(node* b) { (void)b; }
static void n_ssb_destroy(node* b) { n_ssb* s = (n_ssb*)b; free(s->s.p); free(s->sc.p); free(b); }
void* orc_ssb_create(int mode, double bw, double sr, double attack, double decay) {
    NODE_ALLOC(n_ssb);
    n->base.process = n_ssb_proc; n->base.reset = n_ssb_reset; n->base.destroy = n_ssb_destroy;
    double tr = (mode == 0) ? bw / 2.0 : ((mode == 1) ? -bw / 2.0 : 0.0);
    xl_init(&n->x, hz_to_rads(tr, sr));
    agc_init(&n->agc, 1.0, attack, decay, 10e6, 10.0, INFINITY);
    return n;
}

typedef struct { node base; float rate; cf32 off; } n_dc;
static int n_dc_proc(node* b, int c, const void* i, void* o) { n_dc* d = (n_dc*)b; return dcblock_c(d->rate, &d->off, c, (const cf32*)i, (cf32*)o); }
static void n_dc_reset(node* b) { n_dc* d = (n_dc*)b; d->off.re = d->off.im = 0.0f; }
void* orc_dcblock_c_create(double rate) {
    NODE_ALLOC(n_dc);
    n->base.process = n_dc_proc; n->base.reset = n_dc_reset; n->base.destroy = n_plain_destroy;
    n->rate = (float)rate;
    return n;
}

/* ---- noise_reduction::PowerSquelch  (core/src/dsp/noise_reduction/power_squelch.h:33-50): mean amplitude of the chunk
 *      (volk_32fc_magnitude_32f + volk_32f_accumulator_s32f, sequential fp32 sum) against the level in dB ---- */
typedef struct { node base; float level; } n_sq;
static int n_sq_proc(node* b, int count, const void* in, void* out) {
    const cf32* x = (const cf32*)in;
    float sum = 0.0f;
    for (int i = 0; i < count; i++) { sum += sqrtf(x[i].re * x[i].re + x[i].im * x[i].im); }
    sum /= (float)count;
    if (10.0f * log10f(sum) >= ((n_sq*)b)->level) { memcpy(out, in, sizeof(cf32) * (size_t)count); }
    else { memset(out, 0, sizeof(cf32) * (size_t)count); }
    return count;
}
static void n_sq_reset(node* b) { (void)b; }
void* orc_squelch_create(double level) {
    NODE_ALLOC(n_sq);
    n->base.process = n_sq_proc; n->base.reset = n_sq_reset; n->base.destroy = n_plain_destroy;
    n->level = (float)level;
    return n;
}

/* ---- noise_reduction::NoiseBlanker  (core/src/dsp/noise_reduction/noise_blanker.h:12-17,38-57): running mean amplitude
 *      (amp starts at 1), samples more than `level` times above it are scaled back onto it ---- */
typedef struct { node base; float rate, inv_rate, level, amp; } n_nb;
static int n_nb_proc(node* b, int count, const void* in, void* out) {
    n_nb* d = (n_nb*)b;
    const cf32* x = (const cf32*)in;
    cf32* y = (cf32*)out;
    for (int i = 0; i < count; i++) {
        float inAmp = sqrtf((x[i].re * x[i].re) + (x[i].im * x[i].im));      /* complex_t::amplitude (types.h:79-81) */
        float gain = 1.0f;
        if (inAmp != 0.0f) {
            d->amp = (d->amp * d->inv_rate) + (inAmp * d->rate);
            float excess = inAmp / d->amp;
            if (excess > d->level) { gain = 1.0f / excess; }
        }
        y[i].re = x[i].re * gain;
        y[i].im = x[i].im * gain;
    }
    return count;
}
static void n_nb_reset(node* b) { ((n_nb*)b)->amp = 1.0f; }
void* orc_nb_create(double rate, double level) {
    NODE_ALLOC(n_nb);
    n->base.process = n_nb_proc; n->base.reset = n_nb_reset; n->base.destroy = n_plain_destroy;
    n->rate = (float)rate;
    n->inv_rate = 1.0f - n->rate;
    n->level = (float)level;
    n->amp = 1.0f;
    return n;
}

/* ---- noise_reduction::FMIF  (core/src/dsp/noise_reduction/fm_if.h:44-77,95-123): per output sample a Nuttall-windowed
 *      `bins`-point transform of the last `bins` samples, the strongest bin alone transformed back, element bins/2 kept ---- */
typedef struct { node base; int bins; offt_plan* plan; cf32* buffer; size_t cap; float* win; offt_c* fin; offt_c* fout; offt_c* bin; offt_c* bout; float* ampbuf; } n_fmif;
static int n_fmif_proc(node* b, int count, const void* in, void* out) {
    n_fmif* d = (n_fmif*)b;
    const int N = d->bins;
    cf32* y = (cf32*)out;
    if ((size_t)(N - 1 + count) > d->cap) {
        d->cap = (size_t)(N - 1 + count) + 4096;
        d->buffer = (cf32*)realloc(d->buffer, d->cap * sizeof(cf32));
    }
    memcpy(d->buffer + (N - 1), in, sizeof(cf32) * (size_t)count);
    for (int i = 0; i < count; i++) {
        ovk_mul_32fc_32f((ovk_cf32*)d->fin, (const ovk_cf32*)&d->buffer[i], d->win, (unsigned)N);
        offt_forward(d->plan, d->fin, d->fout);
        ovk_magnitude(d->ampbuf, (const ovk_cf32*)d->fout, (unsigned)N);
        unsigned idx = ovk_index_max(d->ampbuf, (unsigned)N);
        d->bin[idx] = d->fout[idx];
        offt_backward(d->plan, d->bin, d->bout);
        y[i].re = d->bout[N / 2].re;
        y[i].im = d->bout[N / 2].im;
        d->bin[idx].re = 0.0f; d->bin[idx].im = 0.0f;
    }
    memmove(d->buffer, d->buffer + count, sizeof(cf32) * (size_t)(N - 1));
    return count;
}
static void n_fmif_reset(node* b) { n_fmif* d = (n_fmif*)b; memset(d->buffer, 0, sizeof(cf32) * (size_t)(d->bins - 1)); }
static void n_fmif_destroy(node* b) {
    n_fmif* d = (n_fmif*)b;
    offt_destroy(d->plan);
    free(d->buffer); free(d->win); free(d->fin); free(d->fout); free(d->bin); free(d->bout); free(d->ampbuf);
    free(d);
}
void* orc_fmif_create(int bins) {
    if (bins < 2) { return NULL; }
    NODE_ALLOC(n_fmif);
    n->base.process = n_fmif_proc; n->base.reset = n_fmif_reset; n->base.destroy = n_fmif_destroy;
    n->bins = bins;
    n->plan = offt_create(bins);
    n->cap = (size_t)bins + 4096;
    n->buffer = (cf32*)calloc(n->cap, sizeof(cf32));
    n->win = (float*)malloc(sizeof(float) * (size_t)bins);
    for (int i = 0; i < bins; i++) { n->win[i] = (float)win_nuttall(i, bins - 1); }       /* fm_if.h:116 */
    n->fin = (offt_c*)calloc((size_t)bins, sizeof(offt_c));
    n->fout = (offt_c*)calloc((size_t)bins, sizeof(offt_c));
    n->bin = (offt_c*)calloc((size_t)bins, sizeof(offt_c));
    n->bout = (offt_c*)calloc((size_t)bins, sizeof(offt_c));
    n->ampbuf = (float*)calloc((size_t)bins, sizeof(float));
    return n;
}

/* ---- filter::Deemphasis<stereo_t>  (core/src/dsp/filter/deephasis.h:14-28,58-77,91-94) ---- */
typedef struct { node base; float alpha; float lastL, lastR; } n_de;
static int n_de_proc(node* b, int count, const void* in, void* out) {
    n_de* d = (n_de*)b;
    const float* x = (const float*)in;
    float* y = (float*)out;
    if (count <= 0) { return count; }
    float a = d->alpha;
    y[0] = (a * x[0]) + ((1 - a) * d->lastL);
    y[1] = (a * x[1]) + ((1 - a) * d->lastR);
    for (int i = 1; i < count; i++) {
        y[2 * i] = (a * x[2 * i]) + ((1 - a) * y[2 * (i - 1)]);
        y[2 * i + 1] = (a * x[2 * i + 1]) + ((1 - a) * y[2 * (i - 1) + 1]);
    }
    d->lastL = y[2 * (count - 1)];
    d->lastR = y[2 * (count - 1) + 1];
    return count;
}
static void n_de_reset(node* b) { n_de* d = (n_de*)b; d->lastL = d->lastR = 0.0f; }
void* orc_deemph_create(double tau, double sr) {
    NODE_ALLOC(n_de);
    n->base.process = n_de_proc; n->base.reset = n_de_reset; n->base.destroy = n_plain_destroy;
    float dt = 1.0f / sr;
    n->alpha = dt / (tau + dt);
    return n;
}

/* ---- RDSDemod, the symbol-rate half of the RDS path  (decoder_modules/radio/src/rds_demod.h:20-73) behind BroadcastFM's rdsOut:
 *      loop::FastAGC<complex_t>(1.0, 1e6, 0.1)            fast_agc.h:61-80
 *   -> loop::Costas<2>(0.005)                             costas.h:18-24,28-33 over PhaseControlLoop (phase_control_loop.h:58-85)
 *   -> filter::FIR<complex_t,complex_t>, taps::bandPass<complex_t>(0, 2375, 100, 5000)     fir.h:74-76, band_pass.h:11-26
 *   -> loop::Costas<2>(0.01, phase 0, freq f = hzToRads(2375/2, 5000), limits f -/+ 10 %)
 *   -> convert::ComplexToReal -> clock_recovery::MM<float>(5000/(2375/2), 1e-6, 0.01, 0.01)   mm.h:94-147 (128 x 8 interpolator bank)
 *   -> digital::BinarySlicer -> digital::DifferentialDecoder(2)
 *      Two outputs per recovered symbol: the soft value (MM output) and the differentially decoded bit.
 *      MM's work buffer is not cleared by the reference (mm.h:36-37: buffer::alloc without clear); its first 7 samples are zero here,
 *      which is what a fresh allocation of that size holds, and the reference build used for pinning clears it too. ---- */
typedef struct { float alpha, beta, phase, freq, minFreq, maxFreq; } pcl_t;
static void pcl_advance(pcl_t* p, float err, int clampPhase, float minPhase, float maxPhase) {
    p->freq += p->beta * err;                                            /* PhaseControlLoop::advance */
    if (p->freq > p->maxFreq) { p->freq = p->maxFreq; }
    else if (p->freq < p->minFreq) { p->freq = p->minFreq; }
    p->phase += p->freq + (p->alpha * err);
    if (clampPhase) {
        const float phaseDelta = maxPhase - minPhase;
        while (p->phase > maxPhase) { p->phase -= phaseDelta; }
        while (p->phase < minPhase) { p->phase += phaseDelta; }
    }
}
static void costas2_process(pcl_t* p, int count, const cf32* in, cf32* out) {
    for (int i = 0; i < count; i++) {
        const float x = -p->phase;
        const cf32 ph = { cosf(x), sinf(x) };                            /* math::phasor(-pcl.phase) */
        cf32 v;
        v.re = (in[i].re * ph.re) - (in[i].im * ph.im);                  /* complex_t * complex_t (types.h:23-25) */
        v.im = (in[i].im * ph.re) + (in[i].re * ph.im);
        out[i] = v;
        float err = v.re * v.im;                                         /* errorFunction, ORDER == 2 */
        if (err < -1.0f) { err = -1.0f; }                                /* std::clamp<float>(err, -1, 1) */
        if (err > 1.0f) { err = 1.0f; }
        pcl_advance(p, err, 1, -FL_M_PI, FL_M_PI);
    }
}
static void pcl_from_pll(pcl_t* c, const pll_t* p) {
    c->alpha = p->alpha; c->beta = p->beta; c->phase = p->phase; c->freq = p->freq; c->minFreq = p->minFreq; c->maxFreq = p->maxFreq;
}
#define RDS_INTERP_PHASES 128
#define RDS_INTERP_TAPS 8
typedef struct {
    float gain, setPoint, maxGain, rate, initGain;       /* FastAGC */
    pcl_t c1, c2, c1_init, c2_init;                        /* the two Costas loops */
    fir_t bp;                                              /* complex taps (ntaps complex, stored as 2*ntaps floats) */
    pcl_t mm; float mm_omega;                              /* MM's PhaseControlLoop<float, false> */
    float bank[RDS_INTERP_PHASES][RDS_INTERP_TAPS];
    float* mmbuf;                                          /* STREAM_BUFFER_SIZE + taps */
    int mm_offset; float lastOut;
    uint8_t diff_last;
    cf32 *a, *b; size_t cap;
} rdsdemod_t;
void* orc_rdsdemod_create(void) {
    rdsdemod_t* r = (rdsdemod_t*)calloc(1, sizeof(rdsdemod_t));
    if (!r) { return NULL; }
    r->setPoint = (float)1.0; r->maxGain = (float)1e6; r->rate = (float)0.1; r->initGain = (float)1.0; r->gain = r->initGain;
    pll_t t;
    pll_init(&t, 0.005f, 0.0, 0.0, -FL_M_PI, FL_M_PI);                 /* costas.init(NULL, 0.005f) */
    pcl_from_pll(&r->c1, &t); r->c1_init = r->c1;
    int nt = orc_bandpass_c(0.0, 2375.0, 100.0, 5000.0, 0, NULL, 0);
    float* bt = (float*)malloc(sizeof(float) * 2 * (size_t)nt);
    orc_bandpass_c(0.0, 2375.0, 100.0, 5000.0, 0, bt, nt);
    fir_init(&r->bp, bt, 2 * nt, 1, 2);
    r->bp.ntaps = nt;
    free(bt);
    const double baudfreq = hz_to_rads(2375.0 / 2.0, 5000.0);
    pll_init(&t, 0.01, 0.0, baudfreq, baudfreq - (baudfreq * 0.1), baudfreq + (baudfreq * 0.1));
    pcl_from_pll(&r->c2, &t); r->c2_init = r->c2;
    /* recov.init(NULL, omega, omegaGain 1e-6, muGain 0.01, omegaRelLimit 0.01): pcl.init(muGain, omegaGain, 0, 0, 1, omega, omega(1-l), omega(1+l)) */
    const double omega = 5000.0 / (2375.0 / 2.0), lim = 0.01;
    r->mm.alpha = (float)0.01; r->mm.beta = (float)1e-6; r->mm.phase = 0.0f; r->mm.freq = (float)omega;
    r->mm.minFreq = (float)(omega * (1.0 - lim)); r->mm.maxFreq = (float)(omega * (1.0 + lim));
    r->mm_omega = (float)omega;
    /* generateInterpTaps (mm.h:168-173): windowedSinc<float>(128*8, hzToRads(0.5/128, 1.0), nuttall, norm = 128), phase-reversed bank */
    {
        const int count = RDS_INTERP_PHASES * RDS_INTERP_TAPS;
        const double om = hz_to_rads(0.5 / (double)RDS_INTERP_PHASES, 1.0);
        const double half = (double)count / 2.0;
        const double corr = (double)RDS_INTERP_PHASES * om / DB_M_PI;
        for (int i = 0; i < count; i++) {
            double tt = (double)i - half + 0.5;
            float tap = (float)(sinc_d(tt * om) * win_nuttall(tt - half, count) * corr);
            r->bank[(RDS_INTERP_PHASES - 1) - (i % RDS_INTERP_PHASES)][i / RDS_INTERP_PHASES] = tap;
        }
    }
    r->mmbuf = (float*)calloc((size_t)STREAM_BUFFER_SIZE + RDS_INTERP_TAPS, sizeof(float));
    return r;
}
void orc_rdsdemod_free(void* h) {
    rdsdemod_t* r = (rdsdemod_t*)h;
    if (!r) { return; }
    fir_free(&r->bp); free(r->mmbuf); free(r->a); free(r->b); free(r);
}
void orc_rdsdemod_reset(void* h) {                                       /* RDSDemod::reset (rds_demod.h:52-62) */
    rdsdemod_t* r = (rdsdemod_t*)h;
    r->gain = r->initGain;
    r->c1.phase = r->c1_init.phase; r->c1.freq = r->c1_init.freq;
    memset(r->bp.buffer, 0, sizeof(cf32) * (size_t)(r->bp.ntaps - 1));
    r->c2.phase = r->c2_init.phase; r->c2.freq = r->c2_init.freq;
    r->mm_offset = 0; r->mm.phase = 0.0f; r->mm.freq = r->mm_omega; r->lastOut = 0.0f;     /* MM::reset keeps the work buffer */
    r->diff_last = 0;
}
int orc_rdsdemod_taps(float* bandpass, int cap_bp, float* bank) {        /* test hook: the two tap sets */
    int nt = orc_bandpass_c(0.0, 2375.0, 100.0, 5000.0, 0, bandpass, cap_bp);
    if (bank) {
        rdsdemod_t* r = (rdsdemod_t*)orc_rdsdemod_create();
        memcpy(bank, r->bank, sizeof(r->bank));
        orc_rdsdemod_free(r);
    }
    return nt;
}
int orc_rdsdemod_process(void* h, int count, const float* in_iq, float* soft, uint8_t* hard) {
    rdsdemod_t* r = (rdsdemod_t*)h;
    const cf32* in = (const cf32*)in_iq;
    if (count < 0 || count > STREAM_BUFFER_SIZE) { return -1; }
    if ((size_t)count > r->cap) {
        r->cap = (size_t)count + 1024;
        r->a = (cf32*)realloc(r->a, r->cap * sizeof(cf32));
        r->b = (cf32*)realloc(r->b, r->cap * sizeof(cf32));
    }
    /* FastAGC<complex_t>::process (fast_agc.h:61-80) */
    for (int i = 0; i < count; i++) {
        r->a[i].re = in[i].re * r->gain;
        r->a[i].im = in[i].im * r->gain;
        float amp = camp(r->a[i]);
        r->gain += (r->setPoint - amp) * r->rate;
        if (r->gain > r->maxGain) { r->gain = r->maxGain; }
    }
    costas2_process(&r->c1, count, r->a, r->b);
    fir_process_cc(&r->bp, count, r->b, r->b);
    costas2_process(&r->c2, count, r->b, r->a);
    /* ComplexToReal, then MM<float>::process (mm.h:94-147) in place */
    float* buf = r->mmbuf;
    for (int i = 0; i < count; i++) { buf[RDS_INTERP_TAPS - 1 + i] = r->a[i].re; }
    int outCount = 0;
    while (r->mm_offset < count) {
        float fph = floorf(r->mm.phase * (float)RDS_INTERP_PHASES);
        int phase = (int)fph;                                            /* std::clamp<int>(floorf(..), 0, phaseCount - 1) */
        if (phase < 0) { phase = 0; }
        if (phase > RDS_INTERP_PHASES - 1) { phase = RDS_INTERP_PHASES - 1; }
        float outVal;
        ovk_dot_32f(&outVal, &buf[r->mm_offset], r->bank[phase], RDS_INTERP_TAPS);
        soft[outCount++] = outVal;
        float sl = (r->lastOut > 0.0) ? 1.0 : -1.0, so = (outVal > 0.0) ? 1.0 : -1.0;    /* math::step<float> (step.h:15) */
        float error = (sl * outVal) - (r->lastOut * so);
        r->lastOut = outVal;
        if (error > 1.0f) { error = 1.0f; }
        if (error < -1.0f) { error = -1.0f; }
        pcl_advance(&r->mm, error, 0, 0.0f, 1.0f);
        float delta = floorf(r->mm.phase);
        r->mm_offset += delta;                                           /* int += float: converted through float */
        r->mm.phase -= delta;
    }
    r->mm_offset -= count;
    memmove(buf, &buf[count], (RDS_INTERP_TAPS - 1) * sizeof(float));
    /* BinarySlicer (binary_slicer.h:14-19), DifferentialDecoder(2) (differential_decoder.h:39-44) */
    for (int i = 0; i < outCount; i++) {
        uint8_t bit = soft[i] > 0.0f;
        hard[i] = (uint8_t)((bit - r->diff_last + 2) % 2);
        r->diff_last = bit;
    }
    return outCount;
}

int orc_process(void* h, int count, const void* in, void* out) { return ((node*)h)->process((node*)h, count, in, out); }
void orc_reset(void* h) { ((node*)h)->reset((node*)h); }
void orc_free(void* h) { if (h) { ((node*)h)->destroy((node*)h); } }

/* ------------------------------------------------------------------ */
/* spectrum branch                                                     */
/* ------------------------------------------------------------------ */

/* IQFrontEnd::genReshapeParams  (core/src/signal_path/iq_frontend.h:59-63) */
void orc_fft_params(double sr, int size, double rate, int* skip, int* nz) {
    int fftInterval = (int)round(sr / rate);
    *nz = fftInterval < size ? fftInterval : size;
    *skip = fftInterval - *nz;
}

/* IQFrontEnd::updateFFTPath window build  (iq_frontend.cpp:281-291): double window value times a float
 * sign, product rounded to float on store */
void orc_window_buf(int win, int nz, float* out) {
    for (int i = 0; i < nz; i++) {
        float sign = (i % 2) ? -1.0f : 1.0f;
        if (win == 0) { out[i] = 1.0f * sign; }
        else if (win == 1) { out[i] = (float)(win_blackman(i, nz) * sign); }
        else { out[i] = (float)(win_nuttall(i, nz) * sign); }
    }
}

typedef struct { int size, nz; float* window; offt_c* in; offt_c* out; offt_plan* plan; } fftpath_t;

void* orc_fft_create(int size, int nz, int win) {
    fftpath_t* f = (fftpath_t*)calloc(1, sizeof(fftpath_t));
    f->size = size;
    f->nz = nz;
    f->window = (float*)malloc(sizeof(float) * (size_t)nz);
    orc_window_buf(win, nz, f->window);
    f->in = (offt_c*)calloc((size_t)size, sizeof(offt_c)); /* zero padding [nz,size) cleared once, :301 */
    f->out = (offt_c*)calloc((size_t)size, sizeof(offt_c));
    f->plan = offt_create(size);
    if (!f->plan) { free(f->window); free(f->in); free(f->out); free(f); return NULL; }
    return f;
}

/* IQFrontEnd::handler  (iq_frontend.cpp:248-267) */
int orc_fft_frame(void* h, const float* iq, float* out_db) {
    fftpath_t* f = (fftpath_t*)h;
    ovk_mul_32fc_32f((cf32*)f->in, (const cf32*)iq, f->window, (unsigned)f->nz);
    offt_forward(f->plan, f->in, f->out);
    ovk_power_spectrum(out_db, (const cf32*)f->out, (float)f->size, (unsigned)f->size);
    return f->size;
}
int orc_fft_raw(void* h, const float* iq, float* out_c) {
    fftpath_t* f = (fftpath_t*)h;
    ovk_mul_32fc_32f((cf32*)f->in, (const cf32*)iq, f->window, (unsigned)f->nz);
    offt_forward(f->plan, f->in, f->out);
    memcpy(out_c, f->out, sizeof(offt_c) * (size_t)f->size);
    return f->size;
}
void orc_fft_free(void* h) {
    fftpath_t* f = (fftpath_t*)h;
    if (!f) { return; }
    free(f->window); free(f->in); free(f->out);
    offt_destroy(f->plan);
    free(f);
}

/* doZoom  (core/src/gui/widgets/waterfall.cpp:65-90): fp32 index accumulator, max-reduce */
void orc_zoom(int offset, int width, int inSize, int outSize, const float* in, float* out) {
    if (offset < 0) { offset = 0; }
    if (width > 524288) { width = 524288; }
    float factor = (float)width / (float)outSize;
    float sFactor = ceilf(factor);
    float id = (float)offset;
    for (int i = 0; i < outSize; i++) {
        float maxVal = -INFINITY;
        int sId = (int)id;
        float uFactor = (sId + sFactor > inSize) ? sFactor - ((sId + sFactor) - inSize) : sFactor;
        for (int j = 0; j < uFactor; j++) {
            if (in[sId + j] > maxVal) { maxVal = in[sId + j]; }
        }
        out[i] = maxVal;
        id += factor;
    }
}

/* FFT hold  (waterfall.cpp:935-939): starts at i = 1 */
void orc_hold(float* hold, const float* latest, int n, float speed) {
    for (int i = 1; i < n; i++) {
        float d = hold[i] - speed;
        hold[i] = (latest[i] < d) ? d : latest[i]; /* std::max<float>(latest, d) */
    }
}

/* file_source int16 ingest  (source_modules/file_source/src/main.cpp:162) */
void orc_i16_to_f32(const int16_t* in, float* out, int n) { ovk_16i_to_32f(out, in, 32768.0f, (unsigned)n); }

/* SampleStreamCompressor::process (dsp/compression/sample_stream_compressor.h:30-66) */
int orc_pcm_compress(int count, int pcmType, const float* iq, uint8_t* out) {
    uint16_t ct = 0, st = (uint16_t)pcmType;
    memcpy(out, &ct, 2);
    memcpy(out + 2, &st, 2);
    if (pcmType == 2) {                                   /* PCM_TYPE_F32: scaler 0, plain copy (:44-48) */
        float z = 0.0f;
        memcpy(out + 4, &z, 4);
        memcpy(out + 8, iq, (size_t)count * 8);
        return 8 + count * 8;
    }
    unsigned int maxIdx = ovk_index_max(iq, (unsigned)count * 2);     /* :51-54: the largest VALUE, not magnitude */
    float maxVal = iq[maxIdx];
    memcpy(out + 4, &maxVal, 4);
    if (pcmType == 0) {                                   /* PCM_TYPE_I8 (:57-59) */
        ovk_32f_to_8i((int8_t*)(out + 8), iq, 128.0f / maxVal, (unsigned)count * 2);
        return 8 + count * 2;
    }
    ovk_32f_to_16i((int16_t*)(out + 8), iq, 32768.0f / maxVal, (unsigned)count * 2);     /* :61-63 */
    return 8 + count * 4;
}

/* SampleStreamDecompressor::process (dsp/compression/sample_stream_decompressor.h:15-37) */
int orc_pcm_decompress(int bytes, const uint8_t* in, float* iq_out) {
    uint16_t st;
    float scaler;
    memcpy(&st, in + 2, 2);
    memcpy(&scaler, in + 4, 4);
    if (st == 2) { memcpy(iq_out, in + 8, (size_t)(bytes - 8)); return (bytes - 8) / 8; }
    if (st == 1) {
        int n = (bytes - 8) / 4;
        ovk_16i_to_32f(iq_out, (const int16_t*)(in + 8), 32768.0f / scaler, (unsigned)n * 2);
        return n;
    }
    if (st == 0) {
        int n = (bytes - 8) / 2;
        ovk_8i_to_32f(iq_out, (const int8_t*)(in + 8), 128.0f / scaler, (unsigned)n * 2);
        return n;
    }
    return 0;
}

/* wav::Writer::write sample conversion (core/src/utils/wav.cpp:150-183) */
void orc_export_convert(const float* in, int n, int type, void* out) {
    if (type == 0) {
        uint8_t* o = (uint8_t*)out;
        for (int i = 0; i < n; i++) { o[i] = (uint8_t)((in[i] * 127.0f) + 128.0f); }          /* :160-163 */
    }
    else if (type == 1) { ovk_32f_to_16i((int16_t*)out, in, 32767.0f, (unsigned)n); }          /* :168 */
    else { ovk_32f_to_32i((int32_t*)out, in, 2147483647.0f, (unsigned)n); }                    /* :172 */
}
