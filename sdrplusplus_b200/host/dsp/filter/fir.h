// host/dsp/filter/fir.h -- dsp::filter::FIR<D,T> with the reference's interface (init / setTaps / reset / process / run,
// core/src/dsp/filter/fir.h:9-96) for the two instantiations on the hot path: complex data x real taps (RxVFO channel
// filter, rx_vfo.h:28-31) and real data x real taps (audio low-pass of the FM demodulators, broadcast_fm.h:45).
// stereo_t data x real taps (the radio module's 300 Hz high-pass, radio_module.h:597-598) is the complex kernel on
// packed float pairs.  The dot products run in libb200dsp (b200_fir_cr_* / b200_fir_rr_*).
#pragma once
#include <type_traits>
#include "../processor.h"
#include "../taps/tap.h"
#include "../b200/handle.h"

namespace dsp::filter {
    template <class D, class T>
    class FIR : public Processor<D, D> {
        using base_type = Processor<D, D>;
        static_assert(std::is_same_v<T, float>, "libb200dsp filters use real taps");
        static_assert(std::is_same_v<D, float> || sizeof(D) == 2 * sizeof(float), "float, complex_t or stereo_t data");
    public:
        FIR() {}
        FIR(stream<D>* in, tap<T>& taps) { init(in, taps); }
        virtual void init(stream<D>* in, tap<T>& taps) {
            build(taps);
            base_type::init(in);
        }
        // new coefficients: the delay line is carried over for the complex kernel (FIR::setTaps, fir.h:31-52);
        // the real kernel starts from a cleared one
        virtual void setTaps(tap<T>& taps) {
            std::lock_guard<std::recursive_mutex> lk(this->ctrlMtx);
            this->tempStop();
            if (!std::is_same_v<D, float> && blk.ok()) { b200_fir_cr_set_taps(blk.get(), taps.taps, (int)taps.size); }
            else { build(taps); }
            this->tempStart();
        }
        virtual void reset() {
            std::lock_guard<std::recursive_mutex> lk(this->ctrlMtx);
            this->tempStop();
            blk.reset();
            this->tempStart();
        }
        bool ok() const { return blk.ok(); }
        inline int process(int count, const D* in, D* out) { return blk.process(count, in, out); }
        DEFAULT_PROC_RUN

    private:
        void build(tap<T>& taps) {
            if constexpr (std::is_same_v<D, float>) { blk.adopt(b200_fir_rr_create(taps.taps, (int)taps.size)); }
            else { blk.adopt(b200_fir_cr_create(taps.taps, (int)taps.size, 1)); }
        }
        b200::Handle blk;
    };
}
