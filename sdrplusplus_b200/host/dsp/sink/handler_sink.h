// host/dsp/sink/handler_sink.h -- dsp::sink::Handler<T> (core/src/dsp/sink/handler_sink.h:6-37): a sink that passes every
// chunk of its input stream to a callback (data, count, ctx) on the block's worker thread.  Runtime glue of the operator API:
// the radio module hangs its RDS group decoder and its symbol display on two of these (demodulators/wfm.h:80-82).
#pragma once
#include "../sink.h"

namespace dsp::sink {
    template <class T>
    class Handler : public Sink<T> {
        using base_type = Sink<T>;
    public:
        typedef void (*callback_t)(T* data, int count, void* ctx);
        Handler() {}
        Handler(stream<T>* in, callback_t handler, void* ctx) { init(in, handler, ctx); }
        void init(stream<T>* in, callback_t handler, void* ctx) {
            cb = handler;
            user = ctx;
            base_type::init(in);
        }
        int run() override {
            const int count = base_type::_in->read();
            if (count < 0) { return -1; }
            if (cb) { cb(base_type::_in->readBuf, count, user); }
            base_type::_in->flush();
            return count;
        }

    private:
        callback_t cb = nullptr;
        void* user = nullptr;
    };
}
