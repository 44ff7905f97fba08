// host/dsp/demod/broadcast_fm.h -- dsp::demod::BroadcastFM with the reference's interface (init(in, deviation,
// samplerate, stereo, lowPass, rdsOut) / setDeviation / setSamplerate / setStereo / setLowPass / setRDSOut / reset /
// process(count, in, out, rdsOutCount, rdsout) / run / rdsOut, core/src/dsp/demod/broadcast_fm.h:20-240), forwarding to
// libb200dsp (b200_wfm_*): discriminator, 19 kHz pilot filter + PLL + L-R recovery (stereo), audio low-pass.
// The RDS side output is the reference's: the demodulated multiplex as complex samples, translated by -57 kHz and resampled
// to 5 kS/s (broadcast_fm.h:52-53,165-170,196-202) -- the stream decoder_modules/radio/src/rds_demod.h reads; it comes from
// its own GPU pass over the same chunk (b200_wfm_rds_create).
#pragma once
#include <vector>
#include "../processor.h"
#include "../b200/handle.h"

namespace dsp::demod {
    class BroadcastFM : public Processor<complex_t, stereo_t> {
        using base_type = Processor<complex_t, stereo_t>;
    public:
        BroadcastFM() {}
        BroadcastFM(stream<complex_t>* in, double deviation, double samplerate, bool stereo = true, bool lowPass = true, bool rdsOut = false) {
            init(in, deviation, samplerate, stereo, lowPass, rdsOut);
        }
        void init(stream<complex_t>* in, double deviation, double samplerate, bool stereo = true, bool lowPass = true, bool rdsOut = false) {
            _dev = deviation; _sr = samplerate; _stereo = stereo; _lowPass = lowPass; _rds = rdsOut;
            make();
            registerOutput(&this->rdsOut);
            base_type::init(in);
        }
        bool ok() const { return blk.ok() && (!_rds || rds.ok()); }
        void setDeviation(double deviation) { _dev = deviation; rebuild(); }
        void setSamplerate(double samplerate) { _sr = samplerate; rebuild(); }
        void setStereo(bool stereo) { _stereo = stereo; rebuild(); }
        void setLowPass(bool lowPass) { _lowPass = lowPass; rebuild(); }
        void setRDSOut(bool rdsOut) { _rds = rdsOut; rebuild(); }
        void reset() {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            tempStop();
            blk.reset();
            rds.reset();
            tempStart();
        }
        // the reference's signature: rdsOutCount / rdsout receive the RDS branch (0 samples while it is off)
        inline int process(int count, complex_t* in, stereo_t* out_, int& rdsOutCount, complex_t* rdsout = nullptr) {
            rdsOutCount = 0;
            if (_rds && rdsout && rds.ok()) {
                const int n = rds.process(count, in, rdsout);
                rdsOutCount = n > 0 ? n : 0;
            }
            return blk.process(count, in, out_);
        }
        inline int process(int count, complex_t* in, stereo_t* out_) {
            int unused = 0;
            return process(count, in, out_, unused, nullptr);
        }
        int run() override {
            const int count = _in->read();
            if (count < 0) { return -1; }
            int rdsCount = 0;
            const int n = process(count, _in->readBuf, out.writeBuf, rdsCount, rdsOut.writeBuf);
            _in->flush();
            if (n < 0) { return -1; }
            if (rdsCount && !rdsOut.swap(rdsCount)) { return -1; }
            if (!out.swap(n)) { return -1; }
            return n;
        }
        stream<complex_t> rdsOut;

    private:
        void make() {
            blk.adopt(b200_wfm_create(_dev, _sr, _stereo ? 1 : 0, _lowPass ? 1 : 0));
            rds.adopt(_rds ? b200_wfm_rds_create(_dev, _sr) : nullptr);
        }
        void rebuild() {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            tempStop();
            make();
            tempStart();
        }
        double _dev = 75000.0, _sr = 250000.0;
        bool _stereo = false, _lowPass = true, _rds = false;
        b200::Handle blk, rds;
    };
}
