// host/dsp/compression/sample_stream_compressor.h -- dsp::compression::SampleStreamCompressor with the reference's
// interface (init(in, pcmType) / setPCMType / static process(count, pcmType, in, out) / run,
// core/src/dsp/compression/sample_stream_compressor.h:5-82), forwarding to libb200dsp (b200_pcm_compress): the
// maximum search and the int8 / int16 conversion of a chunk run on the GPU, the packet comes back byte for byte.
#pragma once
#include "../block.h"

namespace dsp::compression {
    enum PCMType { PCM_TYPE_I8, PCM_TYPE_I16, PCM_TYPE_F32 };       // pcm_type.h

    class SampleStreamCompressor : public Processor<complex_t, uint8_t> {
        using base_type = Processor<complex_t, uint8_t>;
    public:
        SampleStreamCompressor() {}
        SampleStreamCompressor(stream<complex_t>* in, PCMType pcmType) { init(in, pcmType); }
        ~SampleStreamCompressor() override { if (inited) { stop(); } }

        void init(stream<complex_t>* in, PCMType pcmType) {
            _pcmType = pcmType;
            // the reference sizes the output for a full complex buffer + the 8-byte header
            out.setBufferSize(STREAM_BUFFER_SIZE * sizeof(complex_t) + 8);
            base_type::init(in);
        }
        void setPCMType(PCMType pcmType) {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            tempStop();
            _pcmType = pcmType;
            tempStart();
        }
        // returns the packet size in bytes (header + payload), negative on failure
        inline static int process(int count, PCMType pcmType, const complex_t* in, uint8_t* out_) {
            const int fmt = pcmType == PCM_TYPE_F32 ? B200_FMT_CF32 : (pcmType == PCM_TYPE_I16 ? B200_FMT_CS16 : B200_FMT_CS8);
            return b200_pcm_compress(reinterpret_cast<const float*>(in), count, fmt, out_, 8 + count * (int)sizeof(complex_t), B200_MEM_HOST);
        }
        int run() override {
            int count = _in->read();
            if (count < 0) { return -1; }
            int n = process(count, _pcmType, _in->readBuf, out.writeBuf);
            _in->flush();
            if (n < 0) { return -1; }
            if (n && !out.swap(n)) { return -1; }
            return n;
        }
    protected:
        PCMType _pcmType = PCM_TYPE_I16;
    };
}
