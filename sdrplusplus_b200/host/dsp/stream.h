// host/dsp/stream.h -- dsp::stream<T> with the reference's producer/consumer contract
// (writeBuf / readBuf / swap(n) / read() / flush() / stopWriter / stopReader, core/src/dsp/stream.h:25-141),
// written for this library: both buffers are PINNED host memory (b200_host_alloc) so a block's process() can hand
// them to cudaMemcpyAsync without a staging copy, and the size is a constructor / setBufferSize parameter rather
// than a fixed million samples.  One chunk in flight, like the reference.
#pragma once
#include <condition_variable>
#include <mutex>
#include "../../../include/b200dsp.h"

#ifndef STREAM_BUFFER_SIZE
#define STREAM_BUFFER_SIZE 1000000
#endif

namespace dsp {
    class untyped_stream {
    public:
        virtual ~untyped_stream() {}
        virtual bool swap(int) { return false; }
        virtual int read() { return -1; }
        virtual void flush() {}
        virtual void stopWriter() {}
        virtual void clearWriteStop() {}
        virtual void stopReader() {}
        virtual void clearReadStop() {}
    };

    template <class T>
    class stream : public untyped_stream {
    public:
        explicit stream(int samples = STREAM_BUFFER_SIZE) { allocate(samples); }
        ~stream() override { release(); }
        stream(const stream&) = delete;
        stream& operator=(const stream&) = delete;

        void setBufferSize(int samples) { release(); allocate(samples); }
        int bufferSize() const { return cap; }

        // producer: publish `n` samples sitting in writeBuf; blocks while the consumer still owns the other buffer
        bool swap(int n) override {
            std::unique_lock<std::mutex> lk(m);
            cv.wait(lk, [&] { return st == FREE || wstop; });
            if (wstop) { return false; }
            T* t = writeBuf; writeBuf = readBuf; readBuf = t;
            pending = n;
            st = READY;
            cv.notify_all();
            return true;
        }
        // consumer: wait for a published chunk, returns its length or -1 when stopped
        int read() override {
            std::unique_lock<std::mutex> lk(m);
            cv.wait(lk, [&] { return st == READY || rstop; });
            return rstop ? -1 : pending;
        }
        // consumer: done with readBuf
        void flush() override {
            std::lock_guard<std::mutex> lk(m);
            st = FREE;
            cv.notify_all();
        }
        void stopWriter() override { std::lock_guard<std::mutex> lk(m); wstop = true; cv.notify_all(); }
        void clearWriteStop() override { std::lock_guard<std::mutex> lk(m); wstop = false; }
        void stopReader() override { std::lock_guard<std::mutex> lk(m); rstop = true; cv.notify_all(); }
        void clearReadStop() override { std::lock_guard<std::mutex> lk(m); rstop = false; }
        void free() { release(); }

        T* writeBuf = nullptr;
        T* readBuf = nullptr;

    private:
        void allocate(int samples) {
            cap = samples;
            writeBuf = (T*)b200_host_alloc((uint64_t)samples * sizeof(T));
            readBuf = (T*)b200_host_alloc((uint64_t)samples * sizeof(T));
        }
        void release() {
            b200_host_free(writeBuf); b200_host_free(readBuf);
            writeBuf = readBuf = nullptr;
        }
        enum State { FREE, READY } st = FREE;
        std::mutex m;
        std::condition_variable cv;
        int pending = 0, cap = 0;
        bool wstop = false, rstop = false;
    };
}
