// host/dsp/source.h -- dsp::Source<T>: a block that owns its output stream and has no input (core/src/dsp/source.h:6-24)
#pragma once
#include "block.h"

namespace dsp {
    template <class T>
    class Source : public block {
    public:
        Source() { init(); }
        virtual void init() {
            if (inited) { return; }
            registerOutput(&out);
            inited = true;
        }
        stream<T> out;
    };
}
