/*
 * oracle/volk_generic.h  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * From-scratch CPU restatement of the VOLK "generic" (scalar, non-SIMD) leaf
 * kernels that SDR++'s dsp headers call.  VOLK itself is an external, un-pinned
 * dependency of the reference (core/CMakeLists.txt:123-124) and is NOT present
 * in /root/reference nor in this image, so the arithmetic below restates VOLK's
 * published generic-kernel semantics (single sequential fp32 accumulators,
 * rotator renormalised every 512 samples, log2-based power spectrum).
 * PARITY UNPINNED at this boundary: the reference ships no tests or golden
 * vectors for these call sites.
 *
 * Call sites in the reference (what each function stands in for):
 *   ovk_rotator2            core/src/dsp/channel/frequency_xlator.h:45
 *   ovk_dot_32fc_32f        core/src/dsp/filter/decimating_fir.h:56, fir.h:72,
 *                           core/src/dsp/multirate/polyphase_resampler.h:81
 *   ovk_dot_32f             core/src/dsp/filter/fir.h:69
 *   ovk_dot_32fc_32fc       core/src/dsp/filter/fir.h:75
 *   ovk_mul_32fc_32f        core/src/signal_path/iq_frontend.cpp:252
 *   ovk_power_spectrum      core/src/signal_path/iq_frontend.cpp:262
 *   ovk_magnitude           core/src/dsp/demod/am.h:109,120
 *   ovk_16i_to_32f          source_modules/file_source/src/main.cpp:162
 *
 * Build the *declared* oracle with -O2 -ffp-contract=off so that no FMA
 * contraction changes the rounding of these loops.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use anything under oracle/.
 */
#ifndef ORACLE_VOLK_GENERIC_H
#define ORACLE_VOLK_GENERIC_H

#include <math.h>
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; } ovk_cf32;

#define OVK_ROTATOR_RELOAD 512

/* y[i] = x[i]*phase; phase *= inc; |phase| renormalised after every 512
 * samples and once more at the end of the call when the tail was non-empty. */
static inline void ovk_rotator2(ovk_cf32* out, const ovk_cf32* in, const ovk_cf32* inc,
                                ovk_cf32* phase, unsigned int n) {
    unsigned int i, j;
    ovk_cf32 p = *phase;
    const ovk_cf32 d = *inc;
    for (i = 0; i < n / OVK_ROTATOR_RELOAD; ++i) {
        for (j = 0; j < OVK_ROTATOR_RELOAD; ++j) {
            ovk_cf32 x = *in++;
            ovk_cf32 y;
            y.re = x.re * p.re - x.im * p.im;
            y.im = x.re * p.im + x.im * p.re;
            *out++ = y;
            ovk_cf32 q;
            q.re = p.re * d.re - p.im * d.im;
            q.im = p.re * d.im + p.im * d.re;
            p = q;
        }
        float h = hypotf(p.re, p.im);
        p.re /= h;
        p.im /= h;
    }
    for (i = 0; i < n % OVK_ROTATOR_RELOAD; ++i) {
        ovk_cf32 x = *in++;
        ovk_cf32 y;
        y.re = x.re * p.re - x.im * p.im;
        y.im = x.re * p.im + x.im * p.re;
        *out++ = y;
        ovk_cf32 q;
        q.re = p.re * d.re - p.im * d.im;
        q.im = p.re * d.im + p.im * d.re;
        p = q;
    }
    if (i) {
        float h = hypotf(p.re, p.im);
        p.re /= h;
        p.im /= h;
    }
    *phase = p;
}

/* complex data x real taps, one sequential accumulator per component */
static inline void ovk_dot_32fc_32f(ovk_cf32* result, const ovk_cf32* in, const float* taps,
                                    unsigned int n) {
    float re = 0.0f, im = 0.0f;
    for (unsigned int k = 0; k < n; k++) {
        re += in[k].re * taps[k];
        im += in[k].im * taps[k];
    }
    result->re = re;
    result->im = im;
}

static inline void ovk_dot_32f(float* result, const float* in, const float* taps, unsigned int n) {
    float acc = 0.0f;
    for (unsigned int k = 0; k < n; k++) { acc += in[k] * taps[k]; }
    *result = acc;
}

/* complex x complex: two interleaved accumulators (even/odd points), tail added last */
static inline void ovk_dot_32fc_32fc(ovk_cf32* result, const ovk_cf32* in, const ovk_cf32* taps,
                                     unsigned int n) {
    float s0r = 0.0f, s0i = 0.0f, s1r = 0.0f, s1i = 0.0f;
    unsigned int half = n / 2;
    for (unsigned int i = 0; i < half; i++) {
        const ovk_cf32 a0 = in[2 * i], b0 = taps[2 * i];
        const ovk_cf32 a1 = in[2 * i + 1], b1 = taps[2 * i + 1];
        s0r += a0.re * b0.re - a0.im * b0.im;
        s0i += a0.re * b0.im + a0.im * b0.re;
        s1r += a1.re * b1.re - a1.im * b1.im;
        s1i += a1.re * b1.im + a1.im * b1.re;
    }
    float rr = s0r + s1r, ri = s0i + s1i;
    if (n & 1) {
        const ovk_cf32 a = in[n - 1], b = taps[n - 1];
        rr += a.re * b.re - a.im * b.im;
        ri += a.re * b.im + a.im * b.re;
    }
    result->re = rr;
    result->im = ri;
}

static inline void ovk_mul_32fc_32f(ovk_cf32* out, const ovk_cf32* a, const float* b, unsigned int n) {
    for (unsigned int i = 0; i < n; i++) {
        out[i].re = a[i].re * b[i];
        out[i].im = a[i].im * b[i];
    }
}

static inline void ovk_mul_32fc_32fc(ovk_cf32* out, const ovk_cf32* a, const ovk_cf32* b, unsigned int n) {
    for (unsigned int i = 0; i < n; i++) {
        ovk_cf32 x = a[i], y = b[i], r;
        r.re = x.re * y.re - x.im * y.im;
        r.im = x.re * y.im + x.im * y.re;
        out[i] = r;
    }
}

/* log2 with VOLK's non-IEEE floor: log2(0) -> -127 */
static inline float ovk_log2f_non_ieee(float x) {
    float r = log2f(x);
    return isinf(r) ? copysignf(127.0f, r) : r;
}

/* out[k] = 10*log10(|X[k]/norm|^2), VOLK >= 2.x formulation:
 * mag2 -> * 1/norm^2 -> log2 -> * 10/log2(10) */
static inline void ovk_power_spectrum(float* out, const ovk_cf32* in, float norm, unsigned int n) {
    const float normFactSq = 1.0f / (norm * norm);
    const float log2to10 = 3.01029995663981209120f;
    for (unsigned int i = 0; i < n; i++) {
        float m2 = in[i].re * in[i].re + in[i].im * in[i].im;
        m2 = m2 * normFactSq;
        out[i] = ovk_log2f_non_ieee(m2) * log2to10;
    }
}

static inline void ovk_magnitude(float* out, const ovk_cf32* in, unsigned int n) {
    for (unsigned int i = 0; i < n; i++) { out[i] = sqrtf(in[i].re * in[i].re + in[i].im * in[i].im); }
}

static inline void ovk_16i_to_32f(float* out, const int16_t* in, float scalar, unsigned int n) {
    const float inv = 1.0f / scalar;
    for (unsigned int i = 0; i < n; i++) { out[i] = (float)in[i] * inv; }
}

static inline void ovk_8i_to_32f(float* out, const int8_t* in, float scalar, unsigned int n) {
    const float inv = 1.0f / scalar;
    for (unsigned int i = 0; i < n; i++) { out[i] = (float)in[i] * inv; }
}

/* volk_32f_s32f_convert_{8i,16i,32i}: r = in * scalar, clamped to the type's range, rounded with rintf
 * (call sites: sample_stream_compressor.h:56,60; wav.cpp:168,172) */
static inline void ovk_32f_to_8i(int8_t* out, const float* in, float scalar, unsigned int n) {
    for (unsigned int i = 0; i < n; i++) {
        float r = in[i] * scalar;
        if (r > 127.0f) { r = 127.0f; } else if (r < -128.0f) { r = -128.0f; }
        out[i] = (int8_t)rintf(r);
    }
}
static inline void ovk_32f_to_16i(int16_t* out, const float* in, float scalar, unsigned int n) {
    for (unsigned int i = 0; i < n; i++) {
        float r = in[i] * scalar;
        if (r > 32767.0f) { r = 32767.0f; } else if (r < -32768.0f) { r = -32768.0f; }
        out[i] = (int16_t)rintf(r);
    }
}
static inline void ovk_32f_to_32i(int32_t* out, const float* in, float scalar, unsigned int n) {
    for (unsigned int i = 0; i < n; i++) {
        float r = in[i] * scalar;
        if (r > 2147483647.0f) { r = 2147483647.0f; } else if (r < -2147483648.0f) { r = -2147483648.0f; }
        /* (float)INT_MAX is 2^31: the conversion of a clamped top value is out of range on the CPU; pin it */
        out[i] = (r >= 2147483648.0f) ? 2147483647 : (int32_t)rintf(r);
    }
}
/* volk_32f_index_max_32u: first index of the largest value */
static inline unsigned int ovk_index_max(const float* in, unsigned int n) {
    float mx = in[0]; unsigned int idx = 0;
    for (unsigned int i = 1; i < n; i++) { if (in[i] > mx) { mx = in[i]; idx = i; } }
    return idx;
}

#ifdef __cplusplus
}
#endif
#endif
