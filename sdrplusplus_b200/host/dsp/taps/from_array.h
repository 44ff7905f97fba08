// host/dsp/taps/from_array.h -- dsp::taps::fromArray lives with the tap type here (core/src/dsp/taps/from_array.h)
#pragma once
#include "tap.h"
