// host/dsp/multirate/rational_resampler.h -- dsp::multirate::RationalResampler<T> (init / setInSamplerate /
// setOutSamplerate / setRates / reset / process / run, core/src/dsp/multirate/rational_resampler.h:13-179): power-of-two
// pre-decimation + polyphase resampler, planned by the same rule (b200_resamp_plan_get reproduces reconfigure(),
// :120-165) and run on the GPU.  complex_t and stereo_t share the kernels (two packed floats).
#pragma once
#include "../processor.h"
#include "../b200/handle.h"

namespace dsp::multirate {
    template <class T>
    class RationalResampler : public Processor<T, T> {
        using base_type = Processor<T, T>;
        static_assert(sizeof(T) == 2 * sizeof(float), "complex_t or stereo_t");
    public:
        RationalResampler() {}
        RationalResampler(stream<T>* in, double inSamplerate, double outSamplerate) { init(in, inSamplerate, outSamplerate); }
        void init(stream<T>* in, double inSamplerate, double outSamplerate) {
            _inSR = inSamplerate; _outSR = outSamplerate;
            blk.adopt(b200_resamp_create(_inSR, _outSR));
            base_type::init(in);
        }
        void setInSamplerate(double inSamplerate) { _inSR = inSamplerate; rebuild(); }
        void setOutSamplerate(double outSamplerate) { _outSR = outSamplerate; rebuild(); }
        void setRates(double inSamplerate, double outSamplerate) { _inSR = inSamplerate; _outSR = outSamplerate; rebuild(); }
        void reset() {
            std::lock_guard<std::recursive_mutex> lk(this->ctrlMtx);
            this->tempStop();
            blk.reset();
            this->tempStart();
        }
        bool ok() const { return blk.ok(); }
        // upper bound of the output count for `count` inputs (the reference sizes `out` by STREAM_BUFFER_SIZE)
        int maxOut(int count) const { return blk.ok() ? b200_block_max_out(blk.get(), count) : 0; }
        inline int process(int count, const T* in, T* out) { return blk.process(count, in, out); }
        DEFAULT_MULTIRATE_PROC_RUN

    private:
        void rebuild() {
            std::lock_guard<std::recursive_mutex> lk(this->ctrlMtx);
            this->tempStop();
            blk.adopt(b200_resamp_create(_inSR, _outSR));
            this->tempStart();
        }
        double _inSR = 1.0, _outSR = 1.0;
        b200::Handle blk;
    };
}
