// host/dsp/taps/low_pass.h -- dsp::taps::lowPass (core/src/dsp/taps/low_pass.h:7-11): Nuttall-windowed sinc designed in
// fp64 and stored as fp32.  The formula lives once, in libb200dsp (b200_taps_lowpass, bit-identical to the reference:
// tests/test_abi.py), so that the adapter, the library's own blocks and the oracle agree to the last bit.
#pragma once
#include "tap.h"
#include "../../../../include/b200dsp.h"

namespace dsp::taps {
    inline tap<float> lowPass(double cutoff, double transWidth, double samplerate, bool oddTapCount = false) {
        const int n = b200_taps_lowpass(cutoff, transWidth, samplerate, oddTapCount ? 1 : 0, nullptr, 0);
        tap<float> t = alloc<float>(n);
        b200_taps_lowpass(cutoff, transWidth, samplerate, oddTapCount ? 1 : 0, t.taps, n);
        return t;
    }
}
