// GPU-backed stand-in for the reference's sample-stream compressor block.
//
// Interface kept so that the server module binds unchanged (reference: core/src/dsp/compression/
// sample_stream_compressor.h:5-82 -- constructor/init taking the input stream and a PCM type, setPCMType,
// a static process(count, type, in, out) returning the packet size, run()).  The work itself -- maximum search,
// scaling, rounding, saturation, header -- is one call into libb200dsp (b200_pcm_compress); the packet that comes
// back is byte-identical to the CPU block's (tests/test_gpu_parity.py: test_compressed_stream_ingest_and_export).
#pragma once
#include "../block.h"

namespace dsp::compression {

    enum PCMType { PCM_TYPE_I8, PCM_TYPE_I16, PCM_TYPE_F32 };

    namespace detail {
        // PCM type of the packet payload -> sample format code of the C ABI
        inline int abiFormat(PCMType t) {
            switch (t) {
            case PCM_TYPE_F32: return B200_FMT_CF32;
            case PCM_TYPE_I16: return B200_FMT_CS16;
            default: return B200_FMT_CS8;
            }
        }
        constexpr int kHeaderBytes = 8;
    }

    class SampleStreamCompressor : public Processor<complex_t, uint8_t> {
    public:
        SampleStreamCompressor() = default;
        SampleStreamCompressor(stream<complex_t>* source, PCMType type) { init(source, type); }
        ~SampleStreamCompressor() override {
            if (inited) { stop(); }
        }

        void init(stream<complex_t>* source, PCMType type) {
            kind = type;
            // worst case is the float32 payload: a whole complex buffer behind the header
            out.setBufferSize((int)(STREAM_BUFFER_SIZE * sizeof(complex_t)) + detail::kHeaderBytes);
            Processor<complex_t, uint8_t>::init(source);
        }

        void setPCMType(PCMType type) {
            std::lock_guard<std::recursive_mutex> guard(ctrlMtx);
            tempStop();
            kind = type;
            tempStart();
        }

        // Packet size in bytes, or a negative B200_E* code.
        static int process(int count, PCMType type, const complex_t* samples, uint8_t* packet) {
            const int capacity = detail::kHeaderBytes + count * (int)sizeof(complex_t);
            return b200_pcm_compress(reinterpret_cast<const float*>(samples), count, detail::abiFormat(type), packet, capacity, B200_MEM_HOST);
        }

        int run() override {
            const int got = _in->read();
            if (got < 0) { return -1; }
            const int bytes = process(got, kind, _in->readBuf, out.writeBuf);
            _in->flush();
            if (bytes < 0) { return -1; }
            if (bytes > 0 && !out.swap(bytes)) { return -1; }
            return bytes;
        }

    private:
        PCMType kind = PCM_TYPE_I16;
    };
}
