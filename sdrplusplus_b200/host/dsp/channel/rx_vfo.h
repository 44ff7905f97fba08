// host/dsp/channel/rx_vfo.h -- dsp::channel::RxVFO with the reference's public interface
// (init / setInSamplerate / setOutSamplerate / setBandwidth / setOffset / reset / process / run,
// core/src/dsp/channel/rx_vfo.h:17-116); the work is done by libb200dsp (b200_rxvfo_*): translate + decimation
// cascade + polyphase resampler + channel filter in one GPU pass.
#pragma once
#include "../processor.h"
#include "../b200/handle.h"

namespace dsp::channel {
    class RxVFO : public Processor<complex_t, complex_t> {
        using base_type = Processor<complex_t, complex_t>;
    public:
        RxVFO() {}
        RxVFO(stream<complex_t>* in, double inSamplerate, double outSamplerate, double bandwidth, double offset) {
            init(in, inSamplerate, outSamplerate, bandwidth, offset);
        }
        void init(stream<complex_t>* in, double inSamplerate, double outSamplerate, double bandwidth, double offset) {
            _inSR = inSamplerate; _outSR = outSamplerate; _bw = bandwidth; _offset = offset;
            blk.adopt(b200_rxvfo_create(_inSR, _outSR, _bw, _offset));
            base_type::init(in);
        }
        bool ok() const { return blk.ok(); }
        // rate changes rebuild the plan with the worker paused (rx_vfo.h:35-58: tempStop ... tempStart)
        void setInSamplerate(double inSamplerate) { _inSR = inSamplerate; rebuild(); }
        void setOutSamplerate(double outSamplerate, double bandwidth) { _outSR = outSamplerate; _bw = bandwidth; rebuild(); }
        // hot-swappable: applied by the library at the next chunk boundary, phase continuous (rx_vfo.h:60-77)
        void setOffset(double offset) {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            _offset = offset;
            if (blk.ok()) { b200_rxvfo_set_offset(blk.get(), offset); }
        }
        void setBandwidth(double bandwidth) {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            _bw = bandwidth;
            if (blk.ok()) { b200_rxvfo_set_bandwidth(blk.get(), bandwidth); }
        }
        void reset() {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            tempStop();
            blk.reset();
            tempStart();
        }
        // returns the output sample count (0 => nothing to swap), negative on a library error
        inline int process(int count, const complex_t* in, complex_t* out_) { return blk.process(count, in, out_); }
        DEFAULT_MULTIRATE_PROC_RUN

    private:
        void rebuild() {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            tempStop();
            blk.adopt(b200_rxvfo_create(_inSR, _outSR, _bw, _offset));
            tempStart();
        }
        double _inSR = 1.0, _outSR = 1.0, _bw = 1.0, _offset = 0.0;
        b200::Handle blk;
    };
}
