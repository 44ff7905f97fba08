// host/dsp/demod/am.h -- dsp::demod::AM<T> (init / setAGCMode / setBandwidth / setSamplerate / setAGCAttack / setAGCDecay /
// setDCBlockRate / reset / process / run, core/src/dsp/demod/am.h:11-160): [carrier AGC] -> magnitude -> DC blocker ->
// [audio AGC] -> low-pass -> stereo, b200_am_*.  The AGC is the reference's branchy fp32 recurrence run by one GPU thread
// per VFO at the audio rate; its clip look-ahead reaches the end of the chunk, so results depend on the chunking exactly
// as the reference's do (agc.h:93-101).
#pragma once
#include <type_traits>
#include <vector>
#include "../processor.h"
#include "../b200/handle.h"

namespace dsp::demod {
    template <class T>
    class AM : public Processor<complex_t, T> {
        using base_type = Processor<complex_t, T>;
        static_assert(std::is_same_v<T, stereo_t> || std::is_same_v<T, float>, "AM<stereo_t> or AM<float>");
    public:
        enum AGCMode { CARRIER, AUDIO };
        AM() {}
        AM(stream<complex_t>* in, AGCMode agcMode, double bandwidth, double agcAttack, double agcDecay, double dcBlockRate, double samplerate) {
            init(in, agcMode, bandwidth, agcAttack, agcDecay, dcBlockRate, samplerate);
        }
        void init(stream<complex_t>* in, AGCMode agcMode, double bandwidth, double agcAttack, double agcDecay, double dcBlockRate, double samplerate) {
            _mode = agcMode; _bw = bandwidth; _attack = agcAttack; _decay = agcDecay; _dcRate = dcBlockRate; _sr = samplerate;
            blk.adopt(make());
            base_type::init(in);
        }
        void setAGCMode(AGCMode agcMode) { _mode = agcMode; rebuild(); }
        void setBandwidth(double bandwidth) { if (bandwidth != _bw) { _bw = bandwidth; rebuild(); } }
        void setSamplerate(double samplerate) { _sr = samplerate; rebuild(); }
        void setAGCAttack(double attack) { _attack = attack; rebuild(); }
        void setAGCDecay(double decay) { _decay = decay; rebuild(); }
        void setDCBlockRate(double rate) { _dcRate = rate; rebuild(); }
        void reset() {
            std::lock_guard<std::recursive_mutex> lk(this->ctrlMtx);
            this->tempStop();
            blk.reset();
            this->tempStart();
        }
        bool ok() const { return blk.ok(); }
        inline int process(int count, complex_t* in, T* out) {
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
        b200_block* make() const {
            return b200_am_create(_mode == CARRIER ? B200_AGC_CARRIER : B200_AGC_AUDIO, _bw, _attack, _decay, _dcRate, _sr);
        }
        void rebuild() {
            std::lock_guard<std::recursive_mutex> lk(this->ctrlMtx);
            this->tempStop();
            blk.adopt(make());
            this->tempStart();
        }
        AGCMode _mode = AUDIO;
        double _bw = 1.0, _attack = 0.0, _decay = 0.0, _dcRate = 0.0, _sr = 1.0;
        std::vector<stereo_t> lr;
        b200::Handle blk;
    };
}
