// host/dsp/processor.h -- dsp::Processor<I,O> and the run() macros of the operator API (core/src/dsp/processor.h:7-73).
// The macros keep the reference's names and meaning (read -> expression -> flush -> swap); their bodies go through two
// small helpers so that every adapter's run() is the same three calls.
#pragma once
#include "block.h"

namespace dsp {
    template <class I, class O>
    class Processor : public block {
    public:
        Processor() {}
        explicit Processor(stream<I>* in) { init(in); }
        virtual void init(stream<I>* in) {
            _in = in;
            registerInput(_in);
            registerOutput(&out);
            inited = true;
        }
        virtual void setInput(stream<I>* in) {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            tempStop();
            unregisterInput(_in);
            _in = in;
            registerInput(_in);
            tempStart();
        }
        stream<O> out;

    protected:
        // one chunk of a same-rate block: `produced` == consumed; negative = library error -> the worker exits
        int finishChunk(int consumed, int produced, bool alwaysSwap) {
            _in->flush();
            if (produced < 0) { return -1; }
            if ((alwaysSwap || produced > 0) && !out.swap(produced)) { return -1; }
            return consumed;
        }
        stream<I>* _in = nullptr;
    };
}

// int run() whose processing step is `exp` (may use `count`); output count == input count (processor.h:7-19)
#define OVERRIDE_PROC_RUN(exp)                                      \
    int run() override {                                            \
        const int count = base_type::_in->read();                   \
        if (count < 0) { return -1; }                               \
        exp;                                                        \
        return base_type::finishChunk(count, count, true);          \
    }
// `exp` yields the output count; nothing is published when it is 0 (processor.h:21-35)
#define OVERRIDE_MULTIRATE_PROC_RUN(exp)                            \
    int run() override {                                            \
        const int count = base_type::_in->read();                   \
        if (count < 0) { return -1; }                               \
        const int outCount = exp;                                   \
        return base_type::finishChunk(count, outCount, false);      \
    }
#define DEFAULT_PROC_RUN OVERRIDE_PROC_RUN(process(count, base_type::_in->readBuf, base_type::out.writeBuf))
#define DEFAULT_MULTIRATE_PROC_RUN OVERRIDE_MULTIRATE_PROC_RUN(process(count, base_type::_in->readBuf, base_type::out.writeBuf))
