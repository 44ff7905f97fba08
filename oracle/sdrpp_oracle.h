/*
 * oracle/sdrpp_oracle.h  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * One C API, two implementations (loaded into separate ctypes handles):
 *   oracle/liboracle.so          plain-C restatement (sdrpp_oracle.c); travels with the
 *                                repo, needs nothing outside it.
 *   oracle/_ref/libsdrpp_ref.so  the reference's OWN header-only dsp code
 *                                (/root/reference/core/src/dsp, included with -I,
 *                                never copied) wrapped by ref_harness.cpp; built
 *                                only where /root/reference exists.
 * Both sit on the same restated VOLK-generic / FFT leaf layer (volk_generic.h,
 * offt.h): VOLK and FFTW are un-vendored, un-pinned dependencies of the
 * reference and are absent here, and the reference has no tests for this path,
 * so parity is UNPINNED at that leaf boundary.  What IS
 * pinned: the restatement is required to be bit-identical to the reference's
 * own control flow (tests/test_oracle_vs_ref.py) and to the golden fixtures
 * that libsdrpp_ref.so generated (tests/golden/, tools/make_golden.py).
 *
 * Complex samples are interleaved float pairs (re, im) == dsp::complex_t;
 * stereo samples are (l, r) == dsp::stereo_t  (core/src/dsp/types.h:6-127).
 */
#ifndef SDRPP_ORACLE_H
#define SDRPP_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* which implementation is this: "restatement" or "reference-headers" */
const char* orc_impl(void);
/* restatement only: 0 = faithful fp32 rotator recurrence (default), 1 = exact-phase rotator (see sdrpp_oracle.c) */
void orc_set_rotator_mode(int mode);

/* ---- host-side design (taps, windows, plans) ---- */
int orc_estimate_tap_count(double transWidth, double samplerate);
/* taps::lowPass -> returns tap count (writes min(count,cap) taps) */
int orc_lowpass(double cutoff, double transWidth, double samplerate, int oddTapCount, float* out, int cap);
/* taps::highPass (core/src/dsp/taps/high_pass.h:7-14) -> count */
int orc_highpass(double cutoff, double transWidth, double samplerate, int oddTapCount, float* out, int cap);
/* taps::bandPass<complex_t>(bandStart, bandStop, transWidth, sr, odd) -> count (complex taps, interleaved) */
int orc_bandpass_c(double bandStart, double bandStop, double transWidth, double samplerate, int oddTapCount, float* out, int cap);
/* window::{0 rectangular(=1), 1 blackman, 2 nuttall}(n, N) */
double orc_window(int type, double n, double N);
/* decim::plans: returns stage count for a power-of-two ratio (0 if invalid) */
int orc_decim_plan(int ratio, int* decims, int* tapcounts, int cap);
int orc_decim_taps(int ratio, int stage, float* out, int cap);

typedef struct {
    int mode;          /* 0 BOTH, 1 DECIM_ONLY, 2 RESAMP_ONLY, 3 NONE (rational_resampler.h:112-117) */
    int predec_ratio;  /* power-of-two pre-decimation (1 if unused) */
    int interp, decim; /* polyphase L / M */
    int ntaps;         /* prototype tap count */
    int taps_per_phase;
} orc_resamp_plan;
int orc_resamp_plan_get(double inSR, double outSR, orc_resamp_plan* plan);
/* prototype taps already multiplied by interp (rational_resampler.h:159) */
int orc_resamp_taps(double inSR, double outSR, float* out, int cap);

/* ---- streaming blocks: opaque handles, state carried across calls ---- */
void* orc_xlator_create(double offsetHz, double samplerate);             /* FrequencyXlator(in, offset, sr) */
void  orc_xlator_set_offset(void* h, double offsetHz, double samplerate);
void  orc_xlator_get_phase(void* h, float* re_im, float* delta_re_im);
void* orc_decim_create(int ratio);                                       /* PowerDecimator<complex_t> */
void* orc_resamp_create(double inSR, double outSR);                      /* RationalResampler<complex_t> */
void* orc_resamp_stereo_create(double inSR, double outSR);               /* RationalResampler<stereo_t> */
void* orc_fir_cr_create(const float* taps, int n);                       /* FIR<complex_t,float> */
void* orc_fir_rr_create(const float* taps, int n);                       /* FIR<float,float> */
void* orc_decfir_cr_create(const float* taps, int n, int decim);         /* DecimatingFIR<complex_t,float> */
void* orc_rxvfo_create(double inSR, double outSR, double bw, double offset); /* channel::RxVFO */
void  orc_rxvfo_set_offset(void* h, double offset);
void  orc_rxvfo_set_bandwidth(void* h, double bw);
void* orc_quad_create(double deviationHz, double samplerate);            /* demod::Quadrature: cf32 -> f32 */
void* orc_wfm_create(double deviationHz, double samplerate, int stereo, int lowPass); /* BroadcastFM: cf32 -> stereo */
void* orc_wfm_rds_create(double deviationHz, double samplerate);                    /* BroadcastFM's rdsOut: cf32 -> cf32 at 5 kS/s */
void* orc_nfm_create(double samplerate, double bandwidth, int lowPass);  /* FM<stereo_t>: cf32 -> stereo */
void* orc_am_create(int agcMode, double bandwidth, double agcAttack, double agcDecay, double dcBlockRate,
                    double samplerate);                                  /* AM<stereo_t>; agcMode 0 CARRIER 1 AUDIO 2 NONE */
void* orc_ssb_create(int mode, double bandwidth, double samplerate, double agcAttack, double agcDecay);
                                                                         /* SSB<stereo_t>; mode 0 USB 1 LSB 2 DSB */
void* orc_dcblock_c_create(double rate);                                 /* correction::DCBlocker<complex_t> */
void* orc_nb_create(double rate, double level);                            /* noise_reduction::NoiseBlanker */
void* orc_fmif_create(int bins);                                           /* noise_reduction::FMIF */
void* orc_squelch_create(double level);                                   /* noise_reduction::PowerSquelch */
void* orc_deemph_create(double tau, double samplerate);                  /* filter::Deemphasis<stereo_t> */
/* returns output sample count (samples of the block's output type) */
int   orc_process(void* h, int count, const void* in, void* out);
void  orc_reset(void* h);
void  orc_free(void* h);

/* RDSDemod (decoder_modules/radio/src/rds_demod.h:64-73): cf32 at 5 kS/s -> one soft value and one differentially decoded
 * bit per recovered symbol; returns the symbol count of this call */
void* orc_rdsdemod_create(void);
int   orc_rdsdemod_process(void* h, int count, const float* in_iq, float* soft, uint8_t* hard);
void  orc_rdsdemod_reset(void* h);
void  orc_rdsdemod_free(void* h);
int   orc_rdsdemod_taps(float* bandpass, int cap_bp, float* bank);      /* test hook: band-pass taps (count returned), 128 x 8 interpolator bank */

/* ---- spectrum branch (IQFrontEnd::handler / updateFFTPath, iq_frontend.cpp:248-309) ---- */
void  orc_fft_params(double samplerate, int size, double rate, int* skip, int* nz); /* genReshapeParams */
void  orc_window_buf(int window, int nz, float* out);                    /* window(i,nz) * (-1)^i as float */
void* orc_fft_create(int size, int nz, int window);
int   orc_fft_frame(void* h, const float* iq_nz, float* out_db);         /* nz complex in -> size floats out */
int   orc_fft_raw(void* h, const float* iq_nz, float* out_complex);      /* windowed+padded FFT output (size complex) */
void  orc_fft_free(void* h);
/* waterfall.cpp:65-90 doZoom, :935-939 hold */
void  orc_zoom(int offset, int width, int inSize, int outSize, const float* in, float* out);
void  orc_hold(float* hold, const float* latest, int n, float speed);

/* file_source int16 ingest (source_modules/file_source/src/main.cpp:162) */
void  orc_i16_to_f32(const int16_t* in, float* out, int n);
/* compressed sample stream (dsp/compression/sample_stream_{compressor,decompressor}.h): pcmType 0 int8, 1 int16, 2 float32.
 * compress: count complex in -> packet bytes (returned); decompress: packet of `bytes` -> complex count (returned) */
int   orc_pcm_compress(int count, int pcmType, const float* iq, uint8_t* out);
int   orc_pcm_decompress(int bytes, const uint8_t* in, float* iq_out);
/* recorder sample conversion (core/src/utils/wav.cpp:150-183): type 0 uint8, 1 int16, 2 int32 */
void  orc_export_convert(const float* in, int n, int type, void* out);

#ifdef __cplusplus
}
#endif
#endif
