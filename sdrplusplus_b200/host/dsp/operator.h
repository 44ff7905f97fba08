// host/dsp/operator.h -- dsp::Operator<A,B,O>: two inputs, one output (core/src/dsp/operator.h:6-60)
#pragma once
#include "block.h"

namespace dsp {
    template <class A, class B, class O>
    class Operator : public block {
    public:
        Operator() {}
        Operator(stream<A>* a, stream<B>* b) { init(a, b); }
        virtual void init(stream<A>* a, stream<B>* b) {
            _a = a; _b = b;
            registerInput(_a);
            registerInput(_b);
            registerOutput(&out);
            inited = true;
        }
        virtual void setInputs(stream<A>* a, stream<B>* b) {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            tempStop();
            unregisterInput(_a); unregisterInput(_b);
            _a = a; _b = b;
            registerInput(_a); registerInput(_b);
            tempStart();
        }
        virtual void setInputA(stream<A>* a) { setInputs(a, _b); }
        virtual void setInputB(stream<B>* b) { setInputs(_a, b); }
        stream<O> out;

    protected:
        stream<A>* _a = nullptr;
        stream<B>* _b = nullptr;
    };
}
