// host/dsp/channel/frequency_xlator.h -- dsp::channel::FrequencyXlator (init / setOffset / reset / process / run,
// core/src/dsp/channel/frequency_xlator.h:8-60).  The GPU rotates by the closed form of the same fp32-rounded
// phaseDelta; setOffset is phase continuous like the reference's (:25-33).
#pragma once
#include "../processor.h"
#include "../math/hz_to_rads.h"
#include "../b200/handle.h"

namespace dsp::channel {
    class FrequencyXlator : public Processor<complex_t, complex_t> {
        using base_type = Processor<complex_t, complex_t>;
    public:
        FrequencyXlator() {}
        FrequencyXlator(stream<complex_t>* in, double offset, double samplerate) { init(in, offset, samplerate); }
        void init(stream<complex_t>* in, double offset, double samplerate) {
            blk.adopt(b200_xlator_create(offset, samplerate));
            base_type::init(in);
        }
        void setOffset(double offset, double samplerate) {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            if (blk.ok()) { b200_xlator_set_offset(blk.get(), offset, samplerate); }
        }
        void reset() {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            tempStop();
            blk.reset();
            tempStart();
        }
        bool ok() const { return blk.ok(); }
        inline int process(int count, const complex_t* in, complex_t* out_) { return blk.process(count, in, out_); }
        DEFAULT_PROC_RUN

    private:
        b200::Handle blk;
    };
}
