// host/dsp/convert/mono_to_stereo.h -- dsp::convert::MonoToStereo (core/src/dsp/convert/mono_to_stereo.h:7-25).  Inside
// the fused demodulators the duplication is the store of the last kernel; this stand-alone form is a host loop at the
// audio rate for graphs that still name the block.
#pragma once
#include "../processor.h"

namespace dsp::convert {
    class MonoToStereo : public Processor<float, stereo_t> {
        using base_type = Processor<float, stereo_t>;
    public:
        MonoToStereo() {}
        explicit MonoToStereo(stream<float>* in) { base_type::init(in); }
        static inline int process(int count, const float* in, stereo_t* out_) {
            for (int i = 0; i < count; i++) { out_[i].l = in[i]; out_[i].r = in[i]; }
            return count;
        }
        DEFAULT_PROC_RUN
    };
}
