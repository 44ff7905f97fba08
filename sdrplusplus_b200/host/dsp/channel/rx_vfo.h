// host/dsp/channel/rx_vfo.h -- dsp::channel::RxVFO with the reference's public interface
// (init / setOffset / setBandwidth / reset / process / run, core/src/dsp/channel/rx_vfo.h:17-116); the work is done
// by libb200dsp (b200_rxvfo_*): translate + decimation cascade + polyphase resampler + channel filter on the GPU.
#pragma once
#include "../block.h"

namespace dsp::channel {
    class RxVFO : public Processor<complex_t, complex_t> {
        using base_type = Processor<complex_t, complex_t>;
    public:
        RxVFO() {}
        RxVFO(stream<complex_t>* in, double inSamplerate, double outSamplerate, double bandwidth, double offset) {
            init(in, inSamplerate, outSamplerate, bandwidth, offset);
        }
        ~RxVFO() override {
            if (inited) { stop(); }
            b200_block_destroy(h);
        }
        void init(stream<complex_t>* in, double inSamplerate, double outSamplerate, double bandwidth, double offset) {
            h = b200_rxvfo_create(inSamplerate, outSamplerate, bandwidth, offset);
            base_type::init(in);
        }
        bool ok() const { return h != nullptr; }
        void setOffset(double offset) {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            b200_rxvfo_set_offset(h, offset);
        }
        void setBandwidth(double bandwidth) {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            b200_rxvfo_set_bandwidth(h, bandwidth);
        }
        void reset() {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            tempStop();
            b200_block_reset(h);
            tempStart();
        }
        // returns the output sample count (0 => nothing to swap), negative on a library error
        inline int process(int count, const complex_t* in, complex_t* out_) { return b200_block_process(h, count, in, out_); }
        int run() override {
            int count = _in->read();
            if (count < 0) { return -1; }
            int outCount = process(count, _in->readBuf, out.writeBuf);
            _in->flush();
            if (outCount < 0) { return -1; }
            if (outCount && !out.swap(outCount)) { return -1; }
            return outCount;
        }
    private:
        b200_block* h = nullptr;
    };
}
