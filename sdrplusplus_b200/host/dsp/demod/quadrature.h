// host/dsp/demod/quadrature.h -- dsp::demod::Quadrature, the FM discriminator (init / setDeviation / reset / process /
// run, core/src/dsp/demod/quadrature.h:7-60): out = normalizePhase(atan2f(im, re) - previous) / deviation, on the GPU
// (b200_quad_*), previous phase carried across chunks.
#pragma once
#include "../processor.h"
#include "../b200/handle.h"

namespace dsp::demod {
    class Quadrature : public Processor<complex_t, float> {
        using base_type = Processor<complex_t, float>;
    public:
        Quadrature() {}
        Quadrature(stream<complex_t>* in, double deviation, double samplerate) { init(in, deviation, samplerate); }
        void init(stream<complex_t>* in, double deviation, double samplerate) {
            _dev = deviation; _sr = samplerate;
            blk.adopt(b200_quad_create(_dev, _sr));
            base_type::init(in);
        }
        void setDeviation(double deviation, double samplerate) {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            tempStop();
            _dev = deviation; _sr = samplerate;
            blk.adopt(b200_quad_create(_dev, _sr));
            tempStart();
        }
        void reset() {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            tempStop();
            blk.reset();
            tempStart();
        }
        bool ok() const { return blk.ok(); }
        inline int process(int count, const complex_t* in, float* out_) { return blk.process(count, in, out_); }
        DEFAULT_PROC_RUN

    private:
        double _dev = 1.0, _sr = 1.0;
        b200::Handle blk;
    };
}
