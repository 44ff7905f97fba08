// host/dsp/types.h -- sample PODs of the operator API (layout-compatible with the reference's dsp::complex_t /
// dsp::stereo_t, core/src/dsp/types.h:6-127: two packed floats).  Only what the adapters need.
#pragma once
namespace dsp {
    struct complex_t { float re, im; };
    struct stereo_t { float l, r; };
    static_assert(sizeof(complex_t) == 8 && sizeof(stereo_t) == 8, "sample layout");
}
