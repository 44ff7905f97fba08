// host/dsp/sink.h -- dsp::Sink<T>: a block with one input and no output stream (core/src/dsp/sink.h:6-37)
#pragma once
#include "block.h"

namespace dsp {
    template <class T>
    class Sink : public block {
    public:
        Sink() {}
        explicit Sink(stream<T>* in) { init(in); }
        virtual void init(stream<T>* in) {
            _in = in;
            registerInput(_in);
            inited = true;
        }
        virtual void setInput(stream<T>* in) {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            tempStop();
            unregisterInput(_in);
            _in = in;
            registerInput(_in);
            tempStart();
        }

    protected:
        stream<T>* _in = nullptr;
    };
}
