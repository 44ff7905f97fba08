/*
 * oracle/offt.h  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Own fp32 forward complex FFT standing in for FFTW3f, which the reference
 * links as an external, un-pinned dependency (core/CMakeLists.txt:123) and
 * which is not present in this image.  It replaces the call
 *   fftwf_plan_dft_1d(N, in, out, FFTW_FORWARD, FFTW_ESTIMATE) + fftwf_execute
 * at core/src/signal_path/iq_frontend.cpp:255,298:
 *   X[k] = sum_n x[n] * exp(-2*pi*i*n*k/N), unnormalised, single precision.
 * Algorithm: out-of-place Stockham autosort, radix-4 passes plus one radix-2
 * pass when log2(N) is odd; twiddles are computed in double and stored as
 * float.  N must be a power of two (every FFT size the reference UI offers is,
 * core/src/gui/menus/display.cpp:49-58).
 * Cross-checked against numpy's float64 FFT in tests/test_oracle.py.
 */
#ifndef ORACLE_OFFT_H
#define ORACLE_OFFT_H

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; } offt_c;

typedef struct {
    int n;
    offt_c* tw;   /* tw[k] = exp(-2*pi*i*k/n), k < n */
    offt_c* work; /* n scratch elements */
} offt_plan;

/* n a power of two: radix-4 / radix-2 Stockham passes.  Any other n (the 9 / 15 / 31-bin presets of the FM IF noise reduction,
 * decoder_modules/radio/src/radio_module.h:31-36): the defining sum, out[k] = sum_j in[j] tw[(k j) mod n], j ascending. */
static inline offt_plan* offt_create(int n) {
    if (n < 1) { return NULL; }
    offt_plan* p = (offt_plan*)malloc(sizeof(offt_plan));
    p->n = n;
    p->tw = (offt_c*)malloc(sizeof(offt_c) * (size_t)n);
    p->work = (offt_c*)malloc(sizeof(offt_c) * (size_t)n);
    for (int k = 0; k < n; k++) {
        double a = -2.0 * 3.14159265358979323846 * (double)k / (double)n;
        p->tw[k].re = (float)cos(a);
        p->tw[k].im = (float)sin(a);
    }
    return p;
}

static inline void offt_destroy(offt_plan* p) {
    if (!p) { return; }
    free(p->tw);
    free(p->work);
    free(p);
}

static inline offt_c offt_mul(offt_c a, offt_c b) {
    offt_c r;
    r.re = a.re * b.re - a.im * b.im;
    r.im = a.re * b.im + a.im * b.re;
    return r;
}

/* out may alias in.  Result in natural order. */
static inline void offt_forward(const offt_plan* pl, const offt_c* in, offt_c* out) {
    const int N = pl->n;
    if (N & (N - 1)) {
        offt_c* y = pl->work;
        for (int k = 0; k < N; k++) {
            offt_c acc = { 0.0f, 0.0f };
            for (int j = 0; j < N; j++) {
                const offt_c pr = offt_mul(in[j], pl->tw[(int)(((long long)k * j) % N)]);
                acc.re += pr.re;
                acc.im += pr.im;
            }
            y[k] = acc;
        }
        memcpy(out, y, sizeof(offt_c) * (size_t)N);
        return;
    }
    if (in != out) { memcpy(out, in, sizeof(offt_c) * (size_t)N); }
    if (N == 1) { return; }
    offt_c* x = out;
    offt_c* y = pl->work;
    int n = N, s = 1;
    while (n >= 4) {
        const int n1 = n / 4, n2 = n / 2, n3 = n1 + n2;
        const int tstep = N / n;
        for (int p = 0; p < n1; p++) {
            const offt_c w1 = pl->tw[p * tstep];
            const offt_c w2 = pl->tw[2 * p * tstep];
            const offt_c w3 = pl->tw[3 * p * tstep];
            for (int q = 0; q < s; q++) {
                const offt_c a = x[q + s * (p)];
                const offt_c b = x[q + s * (p + n1)];
                const offt_c c = x[q + s * (p + n2)];
                const offt_c d = x[q + s * (p + n3)];
                offt_c apc = { a.re + c.re, a.im + c.im };
                offt_c amc = { a.re - c.re, a.im - c.im };
                offt_c bpd = { b.re + d.re, b.im + d.im };
                /* j*(b-d) */
                offt_c jbmd = { -(b.im - d.im), b.re - d.re };
                offt_c t0 = { apc.re + bpd.re, apc.im + bpd.im };
                offt_c t1 = { amc.re - jbmd.re, amc.im - jbmd.im };
                offt_c t2 = { apc.re - bpd.re, apc.im - bpd.im };
                offt_c t3 = { amc.re + jbmd.re, amc.im + jbmd.im };
                y[q + s * (4 * p + 0)] = t0;
                y[q + s * (4 * p + 1)] = offt_mul(t1, w1);
                y[q + s * (4 * p + 2)] = offt_mul(t2, w2);
                y[q + s * (4 * p + 3)] = offt_mul(t3, w3);
            }
        }
        offt_c* t = x; x = y; y = t;
        n /= 4;
        s *= 4;
    }
    if (n == 2) {
        for (int q = 0; q < s; q++) {
            const offt_c a = x[q], b = x[q + s];
            y[q].re = a.re + b.re;
            y[q].im = a.im + b.im;
            y[q + s].re = a.re - b.re;
            y[q + s].im = a.im - b.im;
        }
        offt_c* t = x; x = y; y = t;
    }
    if (x != out) { memcpy(out, x, sizeof(offt_c) * (size_t)N); }
}

/* Unnormalised inverse transform as the defining sum with conjugated twiddles, out[m] = sum_k in[k] conj(tw[(k m) mod n]),
 * k ascending (any n; the one caller transforms a spectrum with a single non-zero bin, fm_if.h:62-66).  out must not alias in. */
static inline void offt_backward(const offt_plan* pl, const offt_c* in, offt_c* out) {
    const int N = pl->n;
    for (int m = 0; m < N; m++) {
        offt_c acc = { 0.0f, 0.0f };
        for (int k = 0; k < N; k++) {
            const offt_c w = pl->tw[(int)(((long long)k * m) % N)];
            const offt_c cw = { w.re, -w.im };
            const offt_c pr = offt_mul(in[k], cw);
            acc.re += pr.re;
            acc.im += pr.im;
        }
        out[m] = acc;
    }
}

#ifdef __cplusplus
}
#endif
#endif
