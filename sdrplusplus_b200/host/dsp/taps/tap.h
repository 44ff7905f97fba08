// host/dsp/taps/tap.h -- dsp::tap<T>: a plain (pointer, size) view of a coefficient set plus alloc / free
// (core/src/dsp/taps/tap.h:7-30).  Coefficients live in ordinary host memory: they are uploaded once when a block is built.
#pragma once
#include <cstdlib>
#include <cstring>

namespace dsp {
    template <class T>
    struct tap {
        T* taps = nullptr;
        unsigned int size = 0;
    };
    namespace taps {
        template <class T>
        inline tap<T> alloc(int count) {
            tap<T> t;
            t.size = (unsigned)count;
            t.taps = (T*)std::calloc((size_t)(count > 0 ? count : 1), sizeof(T));
            return t;
        }
        template <class T>
        inline void free(tap<T>& t) {
            std::free(t.taps);
            t.taps = nullptr;
            t.size = 0;
        }
        template <class T>
        inline tap<T> fromArray(int count, const T* src) {
            tap<T> t = alloc<T>(count);
            std::memcpy(t.taps, src, (size_t)count * sizeof(T));
            return t;
        }
    }
}
