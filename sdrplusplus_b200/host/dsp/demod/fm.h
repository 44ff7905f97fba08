// host/dsp/demod/fm.h -- dsp::demod::FM<T>, narrow-band FM (init / setSamplerate / setBandwidth / setLowPass / reset /
// process / run, core/src/dsp/demod/fm.h:11-140): Quadrature(deviation = bandwidth / 2) -> optional low-pass
// lowPass(bw/2, 0.1*bw/2, fs) -> mono duplicated to stereo, one GPU pass (b200_nfm_*).  T = stereo_t is what the radio
// module runs (decoder_modules/radio/src/demodulators/nfm.h:75); T = float takes the left channel of the same pass.
// A bandwidth / rate / low-pass change builds a new block: the reference's own updateFilter() clears the filter's
// delay line at that point too (fm.h:123-125).
#pragma once
#include <type_traits>
#include <vector>
#include "../processor.h"
#include "../b200/handle.h"

namespace dsp::demod {
    template <class T>
    class FM : public Processor<complex_t, T> {
        using base_type = Processor<complex_t, T>;
        static_assert(std::is_same_v<T, stereo_t> || std::is_same_v<T, float>, "FM<stereo_t> or FM<float>");
    public:
        FM() {}
        FM(stream<complex_t>* in, double samplerate, double bandwidth, bool lowPass) { init(in, samplerate, bandwidth, lowPass); }
        void init(stream<complex_t>* in, double samplerate, double bandwidth, bool lowPass) {
            _samplerate = samplerate; _bandwidth = bandwidth; _lowPass = lowPass;
            blk.adopt(b200_nfm_create(_samplerate, _bandwidth, _lowPass ? 1 : 0));
            base_type::init(in);
        }
        void setSamplerate(double samplerate) { _samplerate = samplerate; rebuild(); }
        void setBandwidth(double bandwidth) {
            if (bandwidth == _bandwidth) { return; }
            _bandwidth = bandwidth;
            rebuild();
        }
        void setLowPass(bool lowPass) { _lowPass = lowPass; rebuild(); }
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
        void rebuild() {
            std::lock_guard<std::recursive_mutex> lk(this->ctrlMtx);
            this->tempStop();
            blk.adopt(b200_nfm_create(_samplerate, _bandwidth, _lowPass ? 1 : 0));
            this->tempStart();
        }
        double _samplerate = 1.0, _bandwidth = 1.0;
        bool _lowPass = true;
        std::vector<stereo_t> lr;
        b200::Handle blk;
    };
}
