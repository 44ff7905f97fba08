/*
 * oracle/shim/volk/volk.h  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Declaration-compatible stand-in for <volk/volk.h> so that the reference's
 * own header-only dsp library (/root/reference/core/src/dsp, included read-only
 * with -I, nothing copied) compiles in this image, where VOLK is absent.
 * Every volk_* entry point forwards to the scalar restatement in
 * ../../volk_generic.h; see that file for call sites and the parity note.
 */
#pragma once
#include <complex>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>
#include "../../volk_generic.h"

#define VOLK_VERSION 030100

typedef std::complex<float> lv_32fc_t;
typedef std::complex<int16_t> lv_16sc_t;
#define lv_cmake(r, i) lv_32fc_t((float)(r), (float)(i))
#define lv_creal(x) ((x).real())
#define lv_cimag(x) ((x).imag())

static inline size_t volk_get_alignment() { return 64; }
static inline void* volk_malloc(size_t size, size_t alignment) {
    void* p = NULL;
    if (size == 0) { size = alignment; }
    if (posix_memalign(&p, alignment, size) != 0) { return NULL; }
    return p;
}
static inline void volk_free(void* p) { free(p); }

static inline void volk_32fc_s32fc_x2_rotator2_32fc(lv_32fc_t* out, const lv_32fc_t* in, const lv_32fc_t* inc,
                                                    lv_32fc_t* phase, unsigned int n) {
    ovk_rotator2((ovk_cf32*)out, (const ovk_cf32*)in, (const ovk_cf32*)inc, (ovk_cf32*)phase, n);
}
static inline void volk_32fc_s32fc_x2_rotator_32fc(lv_32fc_t* out, const lv_32fc_t* in, const lv_32fc_t inc,
                                                   lv_32fc_t* phase, unsigned int n) {
    ovk_rotator2((ovk_cf32*)out, (const ovk_cf32*)in, (const ovk_cf32*)&inc, (ovk_cf32*)phase, n);
}
static inline void volk_32fc_32f_dot_prod_32fc(lv_32fc_t* r, const lv_32fc_t* in, const float* taps, unsigned int n) {
    ovk_dot_32fc_32f((ovk_cf32*)r, (const ovk_cf32*)in, taps, n);
}
static inline void volk_32f_x2_dot_prod_32f(float* r, const float* in, const float* taps, unsigned int n) {
    ovk_dot_32f(r, in, taps, n);
}
static inline void volk_32fc_x2_dot_prod_32fc(lv_32fc_t* r, const lv_32fc_t* in, const lv_32fc_t* taps, unsigned int n) {
    ovk_dot_32fc_32fc((ovk_cf32*)r, (const ovk_cf32*)in, (const ovk_cf32*)taps, n);
}
static inline void volk_32fc_32f_multiply_32fc(lv_32fc_t* c, const lv_32fc_t* a, const float* b, unsigned int n) {
    ovk_mul_32fc_32f((ovk_cf32*)c, (const ovk_cf32*)a, b, n);
}
static inline void volk_32fc_x2_multiply_32fc(lv_32fc_t* c, const lv_32fc_t* a, const lv_32fc_t* b, unsigned int n) {
    ovk_mul_32fc_32fc((ovk_cf32*)c, (const ovk_cf32*)a, (const ovk_cf32*)b, n);
}
static inline void volk_32fc_s32f_power_spectrum_32f(float* out, const lv_32fc_t* in, const float norm, unsigned int n) {
    ovk_power_spectrum(out, (const ovk_cf32*)in, norm, n);
}
static inline void volk_32fc_magnitude_32f(float* out, const lv_32fc_t* in, unsigned int n) {
    ovk_magnitude(out, (const ovk_cf32*)in, n);
}
static inline void volk_16i_s32f_convert_32f(float* out, const int16_t* in, const float scalar, unsigned int n) {
    ovk_16i_to_32f(out, in, scalar, n);
}
static inline void volk_8i_s32f_convert_32f(float* out, const int8_t* in, const float scalar, unsigned int n) {
    ovk_8i_to_32f(out, in, scalar, n);
}
static inline void volk_32f_x2_interleave_32fc(lv_32fc_t* out, const float* i, const float* q, unsigned int n) {
    for (unsigned int k = 0; k < n; k++) { out[k] = lv_32fc_t(i[k], q[k]); }
}
static inline void volk_32fc_deinterleave_real_32f(float* out, const lv_32fc_t* in, unsigned int n) {
    for (unsigned int k = 0; k < n; k++) { out[k] = in[k].real(); }
}
static inline void volk_32fc_conjugate_32fc(lv_32fc_t* out, const lv_32fc_t* in, unsigned int n) {
    for (unsigned int k = 0; k < n; k++) { out[k] = lv_32fc_t(in[k].real(), -in[k].imag()); }
}
static inline void volk_32f_s32f_multiply_32f(float* c, const float* a, const float s, unsigned int n) {
    for (unsigned int k = 0; k < n; k++) { c[k] = a[k] * s; }
}
static inline void volk_32f_x2_add_32f(float* c, const float* a, const float* b, unsigned int n) {
    for (unsigned int k = 0; k < n; k++) { c[k] = a[k] + b[k]; }
}
static inline void volk_32f_x2_subtract_32f(float* c, const float* a, const float* b, unsigned int n) {
    for (unsigned int k = 0; k < n; k++) { c[k] = a[k] - b[k]; }
}
static inline void volk_32f_x2_multiply_32f(float* c, const float* a, const float* b, unsigned int n) {
    for (unsigned int k = 0; k < n; k++) { c[k] = a[k] * b[k]; }
}
static inline void volk_32f_index_max_32u(uint32_t* target, const float* src, uint32_t n) {
    float mx = src[0]; uint32_t idx = 0;
    for (uint32_t k = 1; k < n; k++) { if (src[k] > mx) { mx = src[k]; idx = k; } }
    *target = idx;
}
static inline void volk_32f_accumulator_s32f(float* result, const float* in, unsigned int n) {
    float acc = 0.0f;
    for (unsigned int k = 0; k < n; k++) { acc += in[k]; }
    *result = acc;
}
static inline void volk_32f_s32f_convert_16i(int16_t* out, const float* in, const float scalar, unsigned int n) {
    for (unsigned int k = 0; k < n; k++) {
        float r = in[k] * scalar;
        if (r > 32767.0f) { r = 32767.0f; } else if (r < -32768.0f) { r = -32768.0f; }
        out[k] = (int16_t)rintf(r);
    }
}
static inline void volk_32f_s32f_convert_32i(int32_t* out, const float* in, const float scalar, unsigned int n) {
    ovk_32f_to_32i(out, in, scalar, n);
}
static inline void volk_32f_s32f_convert_8i(int8_t* out, const float* in, const float scalar, unsigned int n) {
    for (unsigned int k = 0; k < n; k++) {
        float r = in[k] * scalar;
        if (r > 127.0f) { r = 127.0f; } else if (r < -128.0f) { r = -128.0f; }
        out[k] = (int8_t)rintf(r);
    }
}
