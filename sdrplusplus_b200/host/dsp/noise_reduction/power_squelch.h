// host/dsp/noise_reduction/power_squelch.h -- dsp::noise_reduction::PowerSquelch (init / setLevel / process / run,
// core/src/dsp/noise_reduction/power_squelch.h:5-66): a chunk passes when 10 log10(mean |x|) >= level, else it is zeroed.
// The mean and the gate run on the GPU (b200_squelch_*).
#pragma once
#include "../processor.h"
#include "../b200/handle.h"

namespace dsp::noise_reduction {
    class PowerSquelch : public Processor<complex_t, complex_t> {
        using base_type = Processor<complex_t, complex_t>;
    public:
        PowerSquelch() {}
        PowerSquelch(stream<complex_t>* in, double level) { init(in, level); }
        void init(stream<complex_t>* in, double level) {
            _level = level;
            blk.adopt(b200_squelch_create(_level));
            base_type::init(in);
        }
        void setLevel(double level) {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            tempStop();
            _level = level;
            blk.adopt(b200_squelch_create(_level));
            tempStart();
        }
        bool ok() const { return blk.ok(); }
        inline int process(int count, const complex_t* in, complex_t* out_) { return blk.process(count, in, out_); }
        DEFAULT_PROC_RUN

    private:
        double _level = -50.0;
        b200::Handle blk;
    };
}
