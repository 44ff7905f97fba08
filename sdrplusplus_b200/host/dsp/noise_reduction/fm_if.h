// host/dsp/noise_reduction/fm_if.h -- dsp::noise_reduction::FMIF (init / setBins / reset / process / run,
// core/src/dsp/noise_reduction/fm_if.h:6-148): per output sample the strongest bin of a Nuttall-windowed `bins`-point transform of
// the last `bins` samples.  One GPU thread per output sample (b200_fmif_create); the delay line lives in the block.
#pragma once
#include "../processor.h"
#include "../b200/handle.h"

namespace dsp::noise_reduction {
    class FMIF : public Processor<complex_t, complex_t> {
        using base_type = Processor<complex_t, complex_t>;
    public:
        FMIF() {}
        FMIF(stream<complex_t>* in, int bins) { init(in, bins); }
        void init(stream<complex_t>* in, int bins) {
            _bins = bins;
            blk.adopt(b200_fmif_create(_bins));
            base_type::init(in);
        }
        // a new bin count starts from an empty delay line, like the reference's initBuffers()
        void setBins(int bins) {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            tempStop();
            _bins = bins;
            blk.adopt(b200_fmif_create(_bins));
            tempStart();
        }
        void reset() {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            tempStop();
            blk.reset();
            tempStart();
        }
        bool ok() const { return blk.ok(); }
        inline int process(int count, const complex_t* in, complex_t* out_) { return blk.process(count, in, out_); }
        DEFAULT_PROC_RUN

    private:
        int _bins = 32;
        b200::Handle blk;
    };
}
