// host/dsp/demod/broadcast_fm.h -- dsp::demod::BroadcastFM (mono branch) with the reference's interface
// (init(in, deviation, samplerate, stereo, lowPass, rdsOut) / process / run, core/src/dsp/demod/broadcast_fm.h:36-215),
// forwarding to libb200dsp (b200_wfm_*).  Stereo / RDS are not built yet: init() reports failure through ok().
#pragma once
#include "../block.h"

namespace dsp::demod {
    class BroadcastFM : public Processor<complex_t, stereo_t> {
        using base_type = Processor<complex_t, stereo_t>;
    public:
        BroadcastFM() {}
        BroadcastFM(stream<complex_t>* in, double deviation, double samplerate, bool stereo = true, bool lowPass = true, bool rdsOut = false) {
            init(in, deviation, samplerate, stereo, lowPass, rdsOut);
        }
        ~BroadcastFM() override {
            if (inited) { stop(); }
            b200_block_destroy(h);
        }
        void init(stream<complex_t>* in, double deviation, double samplerate, bool stereo = true, bool lowPass = true, bool rdsOut = false) {
            (void)rdsOut;
            h = b200_wfm_create(deviation, samplerate, stereo ? 1 : 0, lowPass ? 1 : 0);
            base_type::init(in);
        }
        bool ok() const { return h != nullptr; }
        void reset() {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            tempStop();
            b200_block_reset(h);
            tempStart();
        }
        inline int process(int count, complex_t* in, stereo_t* out_) { return b200_block_process(h, count, in, out_); }
        int run() override {
            int count = _in->read();
            if (count < 0) { return -1; }
            int n = process(count, _in->readBuf, out.writeBuf);
            _in->flush();
            if (n < 0 || !out.swap(n)) { return -1; }
            return n;
        }
    private:
        b200_block* h = nullptr;
    };
}
