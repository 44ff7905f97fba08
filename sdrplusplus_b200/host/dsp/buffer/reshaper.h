// host/dsp/buffer/reshaper.h -- dsp::buffer::Reshaper<T> (core/src/dsp/buffer/reshaper.h:9-140): re-frames a stream into
// blocks of `keep` samples, `skip` samples apart (skip > 0: samples dropped between blocks; skip < 0: consecutive blocks
// overlap by -skip samples, which come back at the head of the next block).  The spectrum branch's framing of the hot path is
// done inside libb200dsp (b200_fe_set_fft); this host-side block is the runtime glue other users of the type need -- the radio
// module frames the soft RDS symbols for its display with it (demodulators/wfm.h:81, keep 4096, skip 39 - 4096).
// Own implementation: one worker thread and a flat staging vector instead of the reference's ring buffer and second thread.
// The reference's quirk of scaling the overlapped head of complex / stereo blocks by 1/10 (reshaper.h:112-117) is kept.
#pragma once
#include <cstring>
#include <type_traits>
#include <vector>
#include "../block.h"

namespace dsp::buffer {
    template <class T>
    class Reshaper : public block {
    public:
        Reshaper() {}
        Reshaper(stream<T>* in, int keep, int skip) { init(in, keep, skip); }
        ~Reshaper() override { if (inited) { stop(); } }

        void init(stream<T>* in, int keep, int skip) {
            _in = in;
            _keep = keep;
            _skip = skip;
            reframe();
            registerInput(_in);
            registerOutput(&out);
            inited = true;
        }
        void setInput(stream<T>* in) {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            tempStop();
            unregisterInput(_in);
            _in = in;
            registerInput(_in);
            tempStart();
        }
        void setKeep(int keep) {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            tempStop();
            _keep = keep;
            reframe();
            tempStart();
        }
        void setSkip(int skip) {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            tempStop();
            _skip = skip;
            reframe();
            tempStart();
        }

        int run() override {
            const int count = _in->read();
            if (count < 0) { return -1; }
            pending.insert(pending.end(), _in->readBuf, _in->readBuf + count);
            _in->flush();
            // a block needs `fresh` new samples behind the `carry` it re-uses from the previous one
            size_t pos = 0;
            while (true) {
                if (to_drop > 0) {                                  // samples between two blocks (skip > 0)
                    const size_t d = std::min(to_drop, pending.size() - pos);
                    pos += d;
                    to_drop -= d;
                    if (to_drop > 0) { break; }
                }
                if (pending.size() - pos < (size_t)fresh) { break; }
                std::memcpy(blockbuf.data() + carry, pending.data() + pos, sizeof(T) * (size_t)fresh);
                pos += (size_t)fresh;
                std::memcpy(out.writeBuf, blockbuf.data(), sizeof(T) * (size_t)_keep);
                if (!out.swap(_keep)) { return -1; }
                if (carry > 0) {
                    std::memmove(blockbuf.data(), blockbuf.data() + fresh, sizeof(T) * (size_t)carry);
                    if constexpr (std::is_same_v<T, complex_t>) {
                        for (int i = 0; i < carry; i++) { blockbuf[i].re /= 10.0f; blockbuf[i].im /= 10.0f; }
                    }
                    else if constexpr (std::is_same_v<T, stereo_t>) {
                        for (int i = 0; i < carry; i++) { blockbuf[i].l /= 10.0f; blockbuf[i].r /= 10.0f; }
                    }
                }
                to_drop = (size_t)(_skip > 0 ? _skip : 0);
            }
            pending.erase(pending.begin(), pending.begin() + (std::ptrdiff_t)pos);
            return count;
        }

        stream<T> out;

    private:
        void reframe() {
            carry = _skip < 0 ? std::min(-_skip, _keep) : 0;
            fresh = _keep - carry;
            if (fresh < 1) { fresh = 1; carry = _keep - 1; }      // a block always consumes something
            blockbuf.assign((size_t)_keep, T{});
            pending.clear();
            to_drop = 0;
        }
        stream<T>* _in = nullptr;
        int _keep = 1, _skip = 0, carry = 0, fresh = 1;
        size_t to_drop = 0;
        std::vector<T> pending, blockbuf;
    };
}
