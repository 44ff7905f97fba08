// host/dsp/multirate/power_decimator.h -- dsp::multirate::PowerDecimator<T> (init / setRatio / reset / process / run,
// core/src/dsp/multirate/power_decimator.h:12-118): the cascade of DecimatingFIR stages of the pre-computed plan for a
// power-of-two ratio (decim/plans.h:124-139), run as one GPU pass.
#pragma once
#include "../processor.h"
#include "../b200/handle.h"

namespace dsp::multirate {
    template <class T>
    class PowerDecimator : public Processor<T, T> {
        using base_type = Processor<T, T>;
        static_assert(sizeof(T) == 2 * sizeof(float), "complex_t");
    public:
        PowerDecimator() {}
        PowerDecimator(stream<T>* in, unsigned int ratio) { init(in, ratio); }
        void init(stream<T>* in, unsigned int ratio) {
            _ratio = ratio;
            blk.adopt(b200_decim_create((int)_ratio));
            base_type::init(in);
        }
        static inline unsigned int getMaxRatio() { return 8192; }                    // power_decimator.h:28-30
        void setRatio(unsigned int ratio) {
            std::lock_guard<std::recursive_mutex> lk(this->ctrlMtx);
            this->tempStop();
            _ratio = ratio;
            blk.adopt(b200_decim_create((int)_ratio));
            this->tempStart();
        }
        void reset() {
            std::lock_guard<std::recursive_mutex> lk(this->ctrlMtx);
            this->tempStop();
            blk.reset();
            this->tempStart();
        }
        bool ok() const { return blk.ok(); }
        inline int process(int count, const T* in, T* out) { return blk.process(count, in, out); }
        DEFAULT_MULTIRATE_PROC_RUN

    private:
        unsigned int _ratio = 1;
        b200::Handle blk;
    };
}
