/*
 * oracle/ref_harness.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Wraps the reference's OWN header-only dsp library behind the C API of
 * sdrpp_oracle.h.  The headers are included read-only from
 * /root/reference/core/src (compiler flag -I, nothing is copied into this
 * repo); only the VOLK / FFTW leaf layer underneath them is the from-scratch
 * stand-in of oracle/shim (both libraries are absent from this image).
 * The two places where the hot path lives in a .cpp that cannot be compiled
 * here (it pulls GUI globals) are restated below, with their file:line:
 *   IQFrontEnd::handler / updateFFTPath   core/src/signal_path/iq_frontend.cpp:248-309
 *   doZoom / FFT hold                     core/src/gui/widgets/waterfall.cpp:65-90,935-939
 * Built by oracle/Makefile into oracle/_ref/libsdrpp_ref.so (git-ignored).
 */
#include <cstdio>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>

#include <dsp/types.h>
#include <dsp/channel/rx_vfo.h>
#include <dsp/channel/frequency_xlator.h>
#include <dsp/multirate/rational_resampler.h>
#include <dsp/multirate/power_decimator.h>
#include <dsp/filter/fir.h>
#include <dsp/filter/decimating_fir.h>
#include <dsp/filter/deephasis.h>
#include <dsp/demod/quadrature.h>
#include <dsp/demod/broadcast_fm.h>
#include <dsp/demod/fm.h>
#include <dsp/demod/am.h>
#include <dsp/demod/ssb.h>
#include <dsp/correction/dc_blocker.h>
#include <dsp/noise_reduction/power_squelch.h>
#include <dsp/noise_reduction/noise_blanker.h>
#include <dsp/noise_reduction/fm_if.h>
#include <dsp/compression/sample_stream_compressor.h>
#include <dsp/compression/sample_stream_decompressor.h>
#include <dsp/taps/low_pass.h>
#include <dsp/taps/band_pass.h>
#include <dsp/taps/high_pass.h>
#include <dsp/taps/from_array.h>
#include <dsp/window/nuttall.h>
#include <dsp/window/blackman.h>
#include <fftw3.h>
// RDSDemod keeps its blocks private and MM its work buffer protected; the harness needs the latter to give the buffer the
// defined start (zeros) the reference leaves to the allocator (mm.h:36-37).  Standard and dsp headers above are already
// included (include guards), so the two keywords are only redefined for rds_demod.h and clock_recovery/mm.h.
#include <dsp/loop/fast_agc.h>
#include <dsp/loop/costas.h>
#include <dsp/convert/complex_to_real.h>
#include <dsp/digital/binary_slicer.h>
#include <dsp/digital/differential_decoder.h>
#define private public
#define protected public
#include <rds_demod.h>
#undef private
#undef protected
// the radio module's RDS group decoder (block synchronisation, error correction, group parsing on the bit stream): CPU logic
// of the reference, outside the restated path -- compiled from where it lies so that tests can show the bits the demodulator
// delivers are the ones the reference's own decoder understands
#include <rds.h>
#include <rds.cpp>

#include "sdrpp_oracle.h"

using namespace dsp;

namespace {
    struct Node {
        virtual ~Node() {}
        virtual int process(int count, const void* in, void* out) = 0;
        virtual void reset() {}
    };

    // The reference blocks take non-const, sometimes in-place, buffers of at most
    // STREAM_BUFFER_SIZE samples (core/src/dsp/stream.h:9); stage through scratch.
    template <class I>
    struct Scratch {
        std::vector<I> buf;
        I* load(const void* in, int count) {
            if ((int)buf.size() < count) { buf.resize(count); }
            memcpy(buf.data(), in, sizeof(I) * (size_t)count);
            return buf.data();
        }
    };

    struct XlatorNode : Node {
        channel::FrequencyXlator b;
        // expose protected state for tests
        struct Peek : channel::FrequencyXlator { using channel::FrequencyXlator::phase; using channel::FrequencyXlator::phaseDelta; };
        XlatorNode(double off, double sr) { b.init(NULL, off, sr); }
        int process(int count, const void* in, void* out) override { return b.process(count, (const complex_t*)in, (complex_t*)out); }
        void reset() override { b.reset(); }
    };

    struct DecimNode : Node {
        multirate::PowerDecimator<complex_t> b;
        DecimNode(int ratio) { b.init(NULL, ratio); }
        int process(int count, const void* in, void* out) override { return b.process(count, (const complex_t*)in, (complex_t*)out); }
        void reset() override { b.reset(); }
    };

    template <class T>
    struct ResampNode : Node {
        multirate::RationalResampler<T> b;
        ResampNode(double inSR, double outSR) { b.init(NULL, inSR, outSR); }
        int process(int count, const void* in, void* out) override { return b.process(count, (const T*)in, (T*)out); }
        void reset() override { b.reset(); }
    };

    template <class D>
    struct FirNode : Node {
        filter::FIR<D, float> b;
        tap<float> t;
        FirNode(const float* taps, int n) { t = taps::fromArray<float>(n, taps); b.init(NULL, t); }
        ~FirNode() { taps::free(t); }
        int process(int count, const void* in, void* out) override { return b.process(count, (const D*)in, (D*)out); }
        void reset() override { b.reset(); }
    };

    struct DecFirNode : Node {
        filter::DecimatingFIR<complex_t, float> b;
        tap<float> t;
        DecFirNode(const float* taps, int n, int d) { t = taps::fromArray<float>(n, taps); b.init(NULL, t, d); }
        ~DecFirNode() { taps::free(t); }
        int process(int count, const void* in, void* out) override { return b.process(count, (const complex_t*)in, (complex_t*)out); }
        void reset() override { b.reset(); }
    };

    struct RxVfoNode : Node {
        channel::RxVFO b;
        RxVfoNode(double inSR, double outSR, double bw, double off) { b.init(NULL, inSR, outSR, bw, off); }
        int process(int count, const void* in, void* out) override { return b.process(count, (const complex_t*)in, (complex_t*)out); }
        void reset() override { b.reset(); }
    };

    struct QuadNode : Node {
        demod::Quadrature b;
        Scratch<complex_t> s;
        QuadNode(double dev, double sr) { b.init(NULL, dev, sr); }
        int process(int count, const void* in, void* out) override { return b.process(count, s.load(in, count), (float*)out); }
        void reset() override { b.reset(); }
    };

    struct WfmNode : Node {
        demod::BroadcastFM b;
        Scratch<complex_t> s;
        WfmNode(double dev, double sr, bool stereo, bool lp) { b.init(NULL, dev, sr, stereo, lp, false); }
        int process(int count, const void* in, void* out) override {
            int rdsCount = 0;
            return b.process(count, s.load(in, count), (stereo_t*)out, rdsCount, NULL);
        }
        void reset() override { b.reset(); }
    };

    // the RDS side output of BroadcastFM (rdsOut = true): the audio is computed too and dropped
    struct WfmRdsNode : Node {
        demod::BroadcastFM b;
        Scratch<complex_t> s;
        std::vector<stereo_t> audio;
        WfmRdsNode(double dev, double sr, bool stereo) { b.init(NULL, dev, sr, stereo, true, true); }
        int process(int count, const void* in, void* out) override {
            if ((int)audio.size() < count) { audio.resize((size_t)count); }
            int rdsCount = 0;
            b.process(count, s.load(in, count), audio.data(), rdsCount, (complex_t*)out);
            return rdsCount;
        }
        void reset() override { b.reset(); }
    };

    struct NfmNode : Node {
        demod::FM<stereo_t> b;
        Scratch<complex_t> s;
        NfmNode(double sr, double bw, bool lp) { b.init(NULL, sr, bw, lp); }
        int process(int count, const void* in, void* out) override { return b.process(count, s.load(in, count), (stereo_t*)out); }
        void reset() override { b.reset(); }
    };

    struct AmNode : Node {
        demod::AM<stereo_t> b;
        Scratch<complex_t> s;
        int mode;
        AmNode(int agcMode, double bw, double att, double dec, double dcr, double sr) : mode(agcMode) {
            // agcMode 2 (no AGC) is not offered by the reference's enum; map to AUDIO is wrong, so reject
            b.init(NULL, agcMode == 0 ? demod::AM<stereo_t>::CARRIER : demod::AM<stereo_t>::AUDIO, bw, att, dec, dcr, sr);
        }
        int process(int count, const void* in, void* out) override { return b.process(count, s.load(in, count), (stereo_t*)out); }
        void reset() override { b.reset(); }
    };

    struct SsbNode : Node {
        demod::SSB<stereo_t> b;
        SsbNode(int mode, double bw, double sr, double att, double dec) {
            b.init(NULL, (demod::SSB<stereo_t>::Mode)mode, bw, sr, att, dec);
        }
        int process(int count, const void* in, void* out) override { return b.process(count, (const complex_t*)in, (stereo_t*)out); }
    };

    struct DcNode : Node {
        correction::DCBlocker<complex_t> b;
        Scratch<complex_t> s;
        DcNode(double rate) { b.init(NULL, rate); }
        int process(int count, const void* in, void* out) override { return b.process(count, s.load(in, count), (complex_t*)out); }
        void reset() override { b.reset(); }
    };

    struct SquelchNode : Node {
        noise_reduction::PowerSquelch b;
        SquelchNode(double level) { b.init(NULL, level); }
        int process(int count, const void* in, void* out) override { return b.process(count, (const complex_t*)in, (complex_t*)out); }
    };

    struct NbNode : Node {
        noise_reduction::NoiseBlanker b;
        Scratch<complex_t> s;
        NbNode(double rate, double level) { b.init(NULL, rate, level); }
        int process(int count, const void* in, void* out) override { return b.process(count, s.load(in, count), (complex_t*)out); }
        void reset() override { b.reset(); }
    };

    struct FmIfNode : Node {
        noise_reduction::FMIF b;
        FmIfNode(int bins) { b.init(NULL, bins); }
        int process(int count, const void* in, void* out) override { return b.process(count, (const complex_t*)in, (complex_t*)out); }
        void reset() override { b.reset(); }
    };

    struct DeemphNode : Node {
        filter::Deemphasis<stereo_t> b;
        DeemphNode(double tau, double sr) { b.init(NULL, tau, sr); }
        int process(int count, const void* in, void* out) override { return b.process(count, (const stereo_t*)in, (stereo_t*)out); }
        void reset() override { b.reset(); }
    };

    // Spectrum branch: restates IQFrontEnd::updateFFTPath + handler (iq_frontend.cpp:248-309)
    struct FftPath {
        int size, nz;
        float* window;
        fftwf_complex* in;
        fftwf_complex* out;
        fftwf_plan plan;
    };

    int copyTaps(tap<float>& t, float* out, int cap) {
        int n = t.size;
        if (out) { memcpy(out, t.taps, sizeof(float) * (size_t)std::min(n, cap)); }
        return n;
    }
}

extern "C" {

const char* orc_impl(void) { return "reference-headers"; }
void orc_set_rotator_mode(int) {}   // the reference is always its own faithful recurrence

int orc_estimate_tap_count(double tw, double sr) { return taps::estimateTapCount(tw, sr); }

int orc_lowpass(double cutoff, double tw, double sr, int odd, float* out, int cap) {
    tap<float> t = taps::lowPass(cutoff, tw, sr, odd != 0);
    int n = copyTaps(t, out, cap);
    taps::free(t);
    return n;
}

int orc_highpass(double cutoff, double tw, double sr, int odd, float* out, int cap) {
    tap<float> t = taps::highPass(cutoff, tw, sr, odd != 0);
    int n = copyTaps(t, out, cap);
    taps::free(t);
    return n;
}

int orc_bandpass_c(double b0, double b1, double tw, double sr, int odd, float* out, int cap) {
    tap<complex_t> t = taps::bandPass<complex_t>(b0, b1, tw, sr, odd != 0);
    int n = t.size;
    if (out) { memcpy(out, t.taps, sizeof(complex_t) * (size_t)std::min(n, cap)); }
    taps::free(t);
    return n;
}

double orc_window(int type, double n, double N) {
    if (type == 1) { return window::blackman(n, N); }
    if (type == 2) { return window::nuttall(n, N); }
    return 1.0;
}

int orc_decim_plan(int ratio, int* decims, int* tapcounts, int cap) {
    if (ratio < 2 || (ratio & (ratio - 1)) || ratio > (1 << multirate::decim::plans_len)) { return 0; }
    int id = (int)log2(ratio) - 1;
    const multirate::decim::plan& p = multirate::decim::plans[id];
    for (unsigned i = 0; i < p.stageCount && (int)i < cap; i++) {
        decims[i] = p.stages[i].decimation;
        tapcounts[i] = p.stages[i].tapcount;
    }
    return p.stageCount;
}

int orc_decim_taps(int ratio, int stage, float* out, int cap) {
    if (ratio < 2 || (ratio & (ratio - 1)) || ratio > (1 << multirate::decim::plans_len)) { return 0; }
    int id = (int)log2(ratio) - 1;
    const multirate::decim::plan& p = multirate::decim::plans[id];
    if (stage < 0 || stage >= (int)p.stageCount) { return 0; }
    int n = p.stages[stage].tapcount;
    if (out) { memcpy(out, p.stages[stage].taps, sizeof(float) * (size_t)std::min(n, cap)); }
    return n;
}

// Peek into RationalResampler's protected plan by deriving from it.
namespace {
    struct ResampPeek : multirate::RationalResampler<complex_t> {
        using base = multirate::RationalResampler<complex_t>;
        void fill(orc_resamp_plan* pl) {
            pl->mode = (int)base::mode;
            pl->predec_ratio = 1;
            pl->interp = pl->decim = 1;
            pl->ntaps = 0;
            pl->taps_per_phase = 0;
        }
        tap<float>& prototype() { return base::rtaps; }
        int modeId() { return (int)base::mode; }
    };
}

int orc_resamp_plan_get(double inSR, double outSR, orc_resamp_plan* pl) {
    // Re-derive with the same arithmetic as RationalResampler::reconfigure
    // (rational_resampler.h:120-165) and cross-check the mode against the
    // reference object itself.
    ResampPeek r;
    r.init(NULL, inSR, outSR);
    int predecPower = std::min<int>(floor(log2(inSR / outSR)), multirate::PowerDecimator<complex_t>::getMaxRatio());
    int predecRatio = std::min<int>(1 << predecPower, multirate::PowerDecimator<complex_t>::getMaxRatio());
    bool useDecim = (inSR > outSR && predecPower > 0);
    double intSR = useDecim ? inSR / (double)predecRatio : inSR;
    int IntSR = round(intSR);
    int OutSR = round(outSR);
    int g = std::gcd(IntSR, OutSR);
    pl->interp = OutSR / g;
    pl->decim = IntSR / g;
    pl->predec_ratio = useDecim ? predecRatio : 1;
    pl->mode = r.modeId();
    if (pl->interp == pl->decim) {
        pl->ntaps = 0;
        pl->taps_per_phase = 0;
    }
    else {
        pl->ntaps = r.prototype().size;
        pl->taps_per_phase = (pl->ntaps + pl->interp - 1) / pl->interp;
    }
    return 0;
}

int orc_resamp_taps(double inSR, double outSR, float* out, int cap) {
    ResampPeek r;
    r.init(NULL, inSR, outSR);
    if (r.modeId() == 1 || r.modeId() == 3) { return 0; }
    return copyTaps(r.prototype(), out, cap);
}

void* orc_xlator_create(double off, double sr) { return new XlatorNode(off, sr); }
void orc_xlator_set_offset(void* h, double off, double sr) { ((XlatorNode*)h)->b.setOffset(off, sr); }
void orc_xlator_get_phase(void* h, float* ph, float* dl) {
    XlatorNode::Peek* p = (XlatorNode::Peek*)&((XlatorNode*)h)->b;
    ph[0] = p->phase.real(); ph[1] = p->phase.imag();
    dl[0] = p->phaseDelta.real(); dl[1] = p->phaseDelta.imag();
}
void* orc_decim_create(int ratio) { return new DecimNode(ratio); }
void* orc_resamp_create(double inSR, double outSR) { return new ResampNode<complex_t>(inSR, outSR); }
void* orc_resamp_stereo_create(double inSR, double outSR) { return new ResampNode<stereo_t>(inSR, outSR); }
void* orc_fir_cr_create(const float* taps, int n) { return new FirNode<complex_t>(taps, n); }
void* orc_fir_rr_create(const float* taps, int n) { return new FirNode<float>(taps, n); }
void* orc_decfir_cr_create(const float* taps, int n, int d) { return new DecFirNode(taps, n, d); }
void* orc_rxvfo_create(double inSR, double outSR, double bw, double off) { return new RxVfoNode(inSR, outSR, bw, off); }
void orc_rxvfo_set_offset(void* h, double off) { ((RxVfoNode*)h)->b.setOffset(off); }
void orc_rxvfo_set_bandwidth(void* h, double bw) { ((RxVfoNode*)h)->b.setBandwidth(bw); }
void* orc_quad_create(double dev, double sr) { return new QuadNode(dev, sr); }
void* orc_wfm_create(double dev, double sr, int stereo, int lp) { return new WfmNode(dev, sr, stereo != 0, lp != 0); }
void* orc_wfm_rds_create(double dev, double sr) { return new WfmRdsNode(dev, sr, false); }
void* orc_wfm_rds_stereo_create(double dev, double sr) { return new WfmRdsNode(dev, sr, true); }
void* orc_nfm_create(double sr, double bw, int lp) { return new NfmNode(sr, bw, lp != 0); }
void* orc_am_create(int agcMode, double bw, double att, double dec, double dcr, double sr) {
    if (agcMode != 0 && agcMode != 1) { return NULL; }
    return new AmNode(agcMode, bw, att, dec, dcr, sr);
}
void* orc_ssb_create(int mode, double bw, double sr, double att, double dec) { return new SsbNode(mode, bw, sr, att, dec); }
void* orc_dcblock_c_create(double rate) { return new DcNode(rate); }
void* orc_squelch_create(double level) { return new SquelchNode(level); }
void* orc_nb_create(double rate, double level) { return new NbNode(rate, level); }
void* orc_fmif_create(int bins) { return bins >= 2 ? new FmIfNode(bins) : nullptr; }
void* orc_deemph_create(double tau, double sr) { return new DeemphNode(tau, sr); }

// RDSDemod (decoder_modules/radio/src/rds_demod.h): the reference's own class, driven through its process()
struct RdsDemodBox {
    RDSDemod d;
    RdsDemodBox() {
        d.init(NULL, false);
        // recov.out.free() has been called by init; the work buffer is what needs a defined content
        memset(d.recov.buffer, 0, sizeof(float) * (size_t)(STREAM_BUFFER_SIZE + d.recov._interpTapCount));
    }
};
void* orc_rdsdemod_create(void) { return new RdsDemodBox(); }
int orc_rdsdemod_process(void* h, int count, const float* in_iq, float* soft, uint8_t* hard) {
    RdsDemodBox* b = (RdsDemodBox*)h;
    std::vector<complex_t> in((size_t)count + 1);
    memcpy(in.data(), in_iq, sizeof(complex_t) * (size_t)count);
    std::vector<float> so((size_t)count + 16);
    std::vector<uint8_t> ha((size_t)count + 16);
    int n = b->d.process(count, in.data(), so.data(), ha.data());
    memcpy(soft, so.data(), sizeof(float) * (size_t)n);
    memcpy(hard, ha.data(), (size_t)n);
    return n;
}
void orc_rdsdemod_reset(void* h) {
    RdsDemodBox* b = (RdsDemodBox*)h;
    b->d.reset();          // RDSDemod::reset (rds_demod.h:52-62); the blocks are not running, tempStop / tempStart do nothing
}
void orc_rdsdemod_free(void* h) { delete (RdsDemodBox*)h; }
// rds::Decoder (decoder_modules/radio/src/rds.h:214-240): bits in, programme identification and service name out
void* orc_rdsdec_create(void) { return new rds::Decoder(); }
void orc_rdsdec_process(void* h, const uint8_t* bits, int count) {
    std::vector<uint8_t> b(bits, bits + count);
    ((rds::Decoder*)h)->process(b.data(), count);
}
int orc_rdsdec_pi(void* h) { rds::Decoder* d = (rds::Decoder*)h; return d->piCodeValid() ? (int)d->getPICode() : -1; }
int orc_rdsdec_ps(void* h, char* out, int cap) {
    rds::Decoder* d = (rds::Decoder*)h;
    if (!d->PSNameValid()) { return -1; }
    std::string s = d->getPSName(false);
    int n = (int)s.size() < cap - 1 ? (int)s.size() : cap - 1;
    memcpy(out, s.data(), (size_t)n);
    out[n] = 0;
    return n;
}
void orc_rdsdec_free(void* h) { delete (rds::Decoder*)h; }
int orc_rdsdemod_taps(float* bandpass, int cap_bp, float* bank) {
    RdsDemodBox b;
    int nt = b.d.taps.size;
    if (bandpass) { memcpy(bandpass, b.d.taps.taps, sizeof(complex_t) * (size_t)(nt < cap_bp ? nt : cap_bp)); }
    if (bank) {
        for (int p = 0; p < b.d.recov.interpBank.phaseCount; p++) {
            memcpy(bank + (size_t)p * b.d.recov.interpBank.tapsPerPhase, b.d.recov.interpBank.phases[p], sizeof(float) * (size_t)b.d.recov.interpBank.tapsPerPhase);
        }
    }
    return nt;
}

int orc_process(void* h, int count, const void* in, void* out) { return ((Node*)h)->process(count, in, out); }
void orc_reset(void* h) { ((Node*)h)->reset(); }
void orc_free(void* h) { delete (Node*)h; }

// IQFrontEnd::genReshapeParams  (core/src/signal_path/iq_frontend.h:59-63)
void orc_fft_params(double sr, int size, double rate, int* skip, int* nz) {
    int fftInterval = round(sr / rate);
    *nz = std::min<int>(fftInterval, size);
    *skip = fftInterval - *nz;
}

// IQFrontEnd::updateFFTPath window build  (iq_frontend.cpp:281-291)
void orc_window_buf(int win, int nz, float* out) {
    if (win == 0) {
        for (int i = 0; i < nz; i++) { out[i] = 1.0f * ((i % 2) ? -1.0f : 1.0f); }
    }
    else if (win == 1) {
        for (int i = 0; i < nz; i++) { out[i] = window::blackman(i, nz) * ((i % 2) ? -1.0f : 1.0f); }
    }
    else {
        for (int i = 0; i < nz; i++) { out[i] = window::nuttall(i, nz) * ((i % 2) ? -1.0f : 1.0f); }
    }
}

void* orc_fft_create(int size, int nz, int win) {
    FftPath* f = new FftPath;
    f->size = size;
    f->nz = nz;
    f->window = buffer::alloc<float>(nz);
    orc_window_buf(win, nz, f->window);
    f->in = (fftwf_complex*)fftwf_malloc(size * sizeof(fftwf_complex));
    f->out = (fftwf_complex*)fftwf_malloc(size * sizeof(fftwf_complex));
    f->plan = fftwf_plan_dft_1d(size, f->in, f->out, FFTW_FORWARD, FFTW_ESTIMATE);
    // iq_frontend.cpp:301 -- zero padding region cleared once
    buffer::clear(f->in, size - nz, nz);
    return f;
}

// IQFrontEnd::handler  (iq_frontend.cpp:248-267)
int orc_fft_frame(void* h, const float* iq, float* out_db) {
    FftPath* f = (FftPath*)h;
    volk_32fc_32f_multiply_32fc((lv_32fc_t*)f->in, (const lv_32fc_t*)iq, f->window, f->nz);
    fftwf_execute(f->plan);
    volk_32fc_s32f_power_spectrum_32f(out_db, (lv_32fc_t*)f->out, f->size, f->size);
    return f->size;
}

int orc_fft_raw(void* h, const float* iq, float* out_c) {
    FftPath* f = (FftPath*)h;
    volk_32fc_32f_multiply_32fc((lv_32fc_t*)f->in, (const lv_32fc_t*)iq, f->window, f->nz);
    fftwf_execute(f->plan);
    memcpy(out_c, f->out, sizeof(fftwf_complex) * (size_t)f->size);
    return f->size;
}

void orc_fft_free(void* h) {
    FftPath* f = (FftPath*)h;
    buffer::free(f->window);
    fftwf_free(f->in);
    fftwf_free(f->out);
    fftwf_destroy_plan(f->plan);
    delete f;
}

// doZoom  (core/src/gui/widgets/waterfall.cpp:65-90) -- restated: that file cannot be
// compiled here (ImGui / GL).
void orc_zoom(int offset, int width, int inSize, int outSize, const float* in, float* out) {
    if (offset < 0) { offset = 0; }
    if (width > 524288) { width = 524288; }
    float factor = (float)width / (float)outSize;
    float sFactor = ceilf(factor);
    float uFactor;
    float id = offset;
    float maxVal;
    int sId;
    for (int i = 0; i < outSize; i++) {
        maxVal = -INFINITY;
        sId = (int)id;
        uFactor = (sId + sFactor > inSize) ? sFactor - ((sId + sFactor) - inSize) : sFactor;
        for (int j = 0; j < uFactor; j++) {
            if (in[sId + j] > maxVal) { maxVal = in[sId + j]; }
        }
        out[i] = maxVal;
        id += factor;
    }
}

// FFT hold  (waterfall.cpp:935-939) -- note the loop starts at i = 1
void orc_hold(float* hold, const float* latest, int n, float speed) {
    for (int i = 1; i < n; i++) { hold[i] = std::max<float>(latest[i], hold[i] - speed); }
}

void orc_i16_to_f32(const int16_t* in, float* out, int n) { volk_16i_s32f_convert_32f(out, in, 32768.0f, n); }

// the reference's own packet code over the shim's leaf kernels
int orc_pcm_compress(int count, int pcmType, const float* iq, uint8_t* out) {
    return dsp::compression::SampleStreamCompressor::process(count, (dsp::compression::PCMType)pcmType, (const dsp::complex_t*)iq, out);
}
int orc_pcm_decompress(int bytes, const uint8_t* in, float* iq_out) {
    dsp::compression::SampleStreamDecompressor d;
    return d.process(bytes, in, (dsp::complex_t*)iq_out);
}
// wav.cpp is not a header (it pulls the riff writer): the three conversions of Writer::write, line for line in intent
void orc_export_convert(const float* in, int n, int type, void* out) {
    if (type == 0) {
        uint8_t* o = (uint8_t*)out;
        for (int i = 0; i < n; i++) { o[i] = (in[i] * 127.0f) + 128.0f; }
    }
    else if (type == 1) { volk_32f_s32f_convert_16i((int16_t*)out, in, 32767.0f, n); }
    else { volk_32f_s32f_convert_32i((int32_t*)out, in, 2147483647.0f, n); }
}

} // extern "C"
