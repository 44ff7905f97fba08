// host/dsp/taps/high_pass.h -- dsp::taps::highPass (core/src/dsp/taps/high_pass.h:7-14), through b200_taps_highpass
#pragma once
#include "tap.h"
#include "../../../../include/b200dsp.h"

namespace dsp::taps {
    inline tap<float> highPass(double cutoff, double transWidth, double samplerate, bool oddTapCount = false) {
        const int n = b200_taps_highpass(cutoff, transWidth, samplerate, oddTapCount ? 1 : 0, nullptr, 0);
        tap<float> t = alloc<float>(n);
        b200_taps_highpass(cutoff, transWidth, samplerate, oddTapCount ? 1 : 0, t.taps, n);
        return t;
    }
}
