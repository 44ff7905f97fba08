// host/dsp/filter/deephasis.h -- dsp::filter::Deemphasis<stereo_t> (file name as in the reference,
// core/src/dsp/filter/deephasis.h:14-96): y = a x + (1 - a) y[-1] per channel, a = dt / (tau + dt) with dt in float.
// Forwarded to b200_deemph_*; bit-exact against the oracle (tests/test_gpu_parity.py).
#pragma once
#include "../processor.h"
#include "../b200/handle.h"

namespace dsp::filter {
    template <class T>
    class Deemphasis : public Processor<T, T> {
        using base_type = Processor<T, T>;
        static_assert(sizeof(T) == 2 * sizeof(float), "stereo_t");
    public:
        Deemphasis() {}
        Deemphasis(stream<T>* in, double tau, double samplerate) { init(in, tau, samplerate); }
        void init(stream<T>* in, double tau, double samplerate) {
            _tau = tau; _samplerate = samplerate;
            blk.adopt(b200_deemph_create(_tau, _samplerate));
            base_type::init(in);
        }
        void setTau(double tau) { _tau = tau; rebuild(); }
        void setSamplerate(double samplerate) { _samplerate = samplerate; rebuild(); }
        void reset() {
            std::lock_guard<std::recursive_mutex> lk(this->ctrlMtx);
            this->tempStop();
            blk.reset();
            this->tempStart();
        }
        bool ok() const { return blk.ok(); }
        inline int process(int count, const T* in, T* out) { return blk.process(count, in, out); }
        DEFAULT_PROC_RUN

    private:
        void rebuild() {
            std::lock_guard<std::recursive_mutex> lk(this->ctrlMtx);
            this->tempStop();
            blk.adopt(b200_deemph_create(_tau, _samplerate));
            this->tempStart();
        }
        double _tau = 50e-6, _samplerate = 48000.0;
        b200::Handle blk;
    };
}
