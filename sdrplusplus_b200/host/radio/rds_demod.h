// host/radio/rds_demod.h -- RDSDemod of the radio module (decoder_modules/radio/src/rds_demod.h:12-99: init / setSoftEnabled /
// reset / process / run, outputs `out` (decoded bits) and `soft`): takes the place of that file in a build of the radio module
// against host/dsp (wfm.h includes it as "../rds_demod.h").  The whole block -- FastAGC, two Costas loops, band-pass, Mueller &
// Mueller clock recovery, slicer, differential decoder -- is one launch of libb200dsp (b200_rds_demod_process); the symbol count
// of a call comes back from the device with the symbols.
#pragma once
#include <dsp/processor.h>
#include <dsp/b200/handle.h>

class RDSDemod : public dsp::Processor<dsp::complex_t, uint8_t> {
    using base_type = dsp::Processor<dsp::complex_t, uint8_t>;
public:
    RDSDemod() {}
    RDSDemod(dsp::stream<dsp::complex_t>* in, bool enableSoft) { init(in, enableSoft); }
    RDSDemod(const RDSDemod&) = delete;
    RDSDemod& operator=(const RDSDemod&) = delete;
    ~RDSDemod() {
        if (base_type::_block_init) { base_type::stop(); }
        b200_rds_demod_destroy(h);
    }

    void init(dsp::stream<dsp::complex_t>* in, bool enableSoft) {
        this->enableSoft = enableSoft;
        b200_rds_demod_destroy(h);
        h = b200_rds_demod_create();           // nullptr without a CUDA device: process() then reports an error, nothing is computed
        base_type::init(in);
    }

    void setSoftEnabled(bool enable) {
        assert(base_type::_block_init);
        std::lock_guard<std::recursive_mutex> lck(base_type::ctrlMtx);
        base_type::tempStop();
        enableSoft = enable;
        base_type::tempStart();
    }

    void reset() {
        assert(base_type::_block_init);
        std::lock_guard<std::recursive_mutex> lck(base_type::ctrlMtx);
        base_type::tempStop();
        if (h) { b200_rds_demod_reset(h); }
        base_type::tempStart();
    }

    bool ok() const { return h != nullptr; }

    inline int process(int count, dsp::complex_t* in, float* softOut, uint8_t* hardOut) {
        return h ? b200_rds_demod_process(h, count, in, softOut, hardOut) : B200_ESTATE;
    }

    int run() {
        int count = base_type::_in->read();
        if (count < 0) { return -1; }

        count = process(count, base_type::_in->readBuf, soft.writeBuf, base_type::out.writeBuf);
        if (count < 0) { base_type::_in->flush(); return -1; }

        base_type::_in->flush();
        if (!base_type::out.swap(count)) { return -1; }
        if (enableSoft) {
            if (!soft.swap(count)) { return -1; }
        }
        return count;
    }

    dsp::stream<float> soft;

private:
    bool enableSoft = false;
    b200_rds_demod* h = nullptr;
};
