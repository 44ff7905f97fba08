// host/dsp/math/hz_to_rads.h -- dsp::math::hzToRads (core/src/dsp/math/hz_to_rads.h:6-8)
#pragma once
namespace dsp::math {
    inline double hzToRads(double freq, double samplerate) { return 2.0 * 3.14159265358979323846 * (freq / samplerate); }
}
