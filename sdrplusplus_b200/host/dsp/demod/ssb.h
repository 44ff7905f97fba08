// host/dsp/demod/ssb.h -- dsp::demod::SSB<T> (init / setMode / setBandwidth / setSamplerate / setAGCAttack / setAGCDecay /
// reset / process / run, core/src/dsp/demod/ssb.h:10-130): second translator by +bw/2 (USB), -bw/2 (LSB) or 0 (DSB)
// (:106-116) -> real part -> AGC -> stereo, b200_ssb_*.  The audio-rate translator is the reference's faithful fp32
// recurrence (one GPU thread per VFO).
#pragma once
#include <type_traits>
#include <vector>
#include "../processor.h"
#include "../b200/handle.h"

namespace dsp::demod {
    template <class T>
    class SSB : public Processor<complex_t, T> {
        using base_type = Processor<complex_t, T>;
        static_assert(std::is_same_v<T, stereo_t> || std::is_same_v<T, float>, "SSB<stereo_t> or SSB<float>");
    public:
        enum Mode { USB, LSB, DSB };
        SSB() {}
        SSB(stream<complex_t>* in, Mode mode, double bandwidth, double samplerate, double agcAttack, double agcDecay) {
            init(in, mode, bandwidth, samplerate, agcAttack, agcDecay);
        }
        void init(stream<complex_t>* in, Mode mode, double bandwidth, double samplerate, double agcAttack, double agcDecay) {
            _mode = mode; _bw = bandwidth; _sr = samplerate; _attack = agcAttack; _decay = agcDecay;
            blk.adopt(make());
            base_type::init(in);
        }
        void setMode(Mode mode) { _mode = mode; rebuild(); }
        void setBandwidth(double bandwidth) { _bw = bandwidth; rebuild(); }
        void setSamplerate(double samplerate) { _sr = samplerate; rebuild(); }
        void setAGCAttack(double attack) { _attack = attack; rebuild(); }
        void setAGCDecay(double decay) { _decay = decay; rebuild(); }
        bool ok() const { return blk.ok(); }
        inline int process(int count, const complex_t* in, T* out) {
            if constexpr (std::is_same_v<T, stereo_t>) { return blk.process(count, in, out); }
            else {
                if ((int)lr.size() < count) { lr.resize((size_t)count); }
                const int n = blk.process(count, in, lr.data());
                for (int i = 0; i < n; i++) { out[i] = lr[(size_t)i].l; }
                return n;
            }
        }
        DEFAULT_PROC_RUN

    private:
        b200_block* make() const { return b200_ssb_create((int)_mode, _bw, _sr, _attack, _decay); }
        void rebuild() {
            std::lock_guard<std::recursive_mutex> lk(this->ctrlMtx);
            this->tempStop();
            blk.adopt(make());
            this->tempStart();
        }
        Mode _mode = USB;
        double _bw = 1.0, _sr = 1.0, _attack = 0.0, _decay = 0.0;
        std::vector<stereo_t> lr;
        b200::Handle blk;
    };
}
