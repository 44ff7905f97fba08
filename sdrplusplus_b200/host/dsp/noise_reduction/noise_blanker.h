// host/dsp/noise_reduction/noise_blanker.h -- dsp::noise_reduction::NoiseBlanker (init / setRate / setLevel / reset / process /
// run, core/src/dsp/noise_reduction/noise_blanker.h:5-77): a running mean of the amplitude; samples more than `level` times above
// it are scaled back onto it.  The recurrence runs on the GPU (b200_noise_blanker_*), one thread per stream like the reference's loop.
#pragma once
#include "../processor.h"
#include "../b200/handle.h"

namespace dsp::noise_reduction {
    class NoiseBlanker : public Processor<complex_t, complex_t> {
        using base_type = Processor<complex_t, complex_t>;
    public:
        NoiseBlanker() {}
        NoiseBlanker(stream<complex_t>* in, double rate, double level) { init(in, rate, level); }
        void init(stream<complex_t>* in, double rate, double level) {
            _rate = rate;
            _level = level;
            blk.adopt(b200_noise_blanker_create(_rate, _level));
            base_type::init(in);
        }
        // both take effect at the next chunk; the running amplitude carries over, as in the reference
        void setRate(double rate) {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            _rate = rate;
            if (blk.ok()) { b200_noise_blanker_set(blk.get(), _rate, _level); }
        }
        void setLevel(double level) {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            _level = level;
            if (blk.ok()) { b200_noise_blanker_set(blk.get(), _rate, _level); }
        }
        void reset() {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            blk.reset();
        }
        bool ok() const { return blk.ok(); }
        inline int process(int count, complex_t* in, complex_t* out_) { return blk.process(count, in, out_); }
        DEFAULT_PROC_RUN

    private:
        double _rate = 0.0, _level = 10.0;
        b200::Handle blk;
    };
}
