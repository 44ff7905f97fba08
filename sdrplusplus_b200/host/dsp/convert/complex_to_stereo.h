// host/dsp/convert/complex_to_stereo.h -- dsp::convert::ComplexToStereo (core/src/dsp/convert/complex_to_stereo.h:7-26):
// (re, im) reinterpreted as (l, r); the radio module's RAW "demodulator" (decoder_modules/radio/src/demodulators/raw.h:68).
#pragma once
#include <cstring>
#include "../processor.h"

namespace dsp::convert {
    class ComplexToStereo : public Processor<complex_t, stereo_t> {
        using base_type = Processor<complex_t, stereo_t>;
    public:
        ComplexToStereo() {}
        explicit ComplexToStereo(stream<complex_t>* in) { base_type::init(in); }
        static inline int process(int count, const complex_t* in, stereo_t* out_) {
            std::memcpy(out_, in, (size_t)count * sizeof(complex_t));
            return count;
        }
        DEFAULT_PROC_RUN
    };
}
