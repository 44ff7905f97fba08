/*
 * oracle/shim/fftw3.h  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Minimal stand-in for the single-precision FFTW3 API used at
 * core/src/signal_path/iq_frontend.cpp:60-62,255,294-298, backed by ../offt.h.
 */
#pragma once
#include <cstdlib>
#include "../offt.h"

typedef float fftwf_complex[2];
struct fftwf_plan_s { offt_plan* p; fftwf_complex* in; fftwf_complex* out; int sign; };
typedef fftwf_plan_s* fftwf_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_ESTIMATE (1U << 6)

static inline void* fftwf_malloc(size_t n) { void* p = NULL; if (posix_memalign(&p, 64, n ? n : 64)) { return NULL; } return p; }
static inline void fftwf_free(void* p) { free(p); }
static inline fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex* in, fftwf_complex* out, int sign, unsigned) {
    fftwf_plan pl = new fftwf_plan_s;
    pl->p = offt_create(n); pl->in = in; pl->out = out; pl->sign = sign;
    return pl;
}
static inline void fftwf_execute(const fftwf_plan pl) {
    if (pl->sign == FFTW_BACKWARD) { offt_backward(pl->p, (const offt_c*)pl->in, (offt_c*)pl->out); }   // core/src/dsp/noise_reduction/fm_if.h:121
    else { offt_forward(pl->p, (const offt_c*)pl->in, (offt_c*)pl->out); }
}
static inline void fftwf_destroy_plan(fftwf_plan pl) { if (pl) { offt_destroy(pl->p); delete pl; } }
