// host/dsp/block.h -- worker-thread block base of the operator API (start / stop / tempStart / tempStop / run,
// core/src/dsp/block.h:18-131), for adapters whose process() forwards to libb200dsp.  One worker thread per started
// block, `while (run() >= 0)`.  The typed bases live in processor.h / sink.h / source.h / operator.h like the reference.
#pragma once
#include <algorithm>
#include <cassert>
#include <mutex>
#include <thread>
#include <vector>
#include "stream.h"
#include "types.h"

namespace dsp {
    // what hier_block and chain hold: anything that can be started and stopped (block.h:10-16)
    class generic_block {
    public:
        virtual ~generic_block() {}
        virtual void start() {}
        virtual void stop() {}
        virtual int run() { return -1; }
    };

    class block : public generic_block {
    public:
        ~block() override { if (inited) { stop(); } }
        void start() override {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            if (running) { return; }
            running = true;
            launch();
        }
        void stop() override {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            if (!running) { return; }
            halt();
            running = false;
        }
        void tempStop() {
            if (depth++ == 0 && running && !paused) { halt(); paused = true; }
        }
        void tempStart() {
            if (depth > 0 && --depth == 0 && paused) { launch(); paused = false; }
        }
        int run() override = 0;

    protected:
        void registerInput(untyped_stream* s) { ins.push_back(s); }
        void unregisterInput(untyped_stream* s) { ins.erase(std::remove(ins.begin(), ins.end(), s), ins.end()); }
        void registerOutput(untyped_stream* s) { outs.push_back(s); }
        void unregisterOutput(untyped_stream* s) { outs.erase(std::remove(outs.begin(), outs.end(), s), outs.end()); }
        bool inited = false;
        bool& _block_init = inited;
        std::recursive_mutex ctrlMtx;

    private:
        void launch() { worker = std::thread([this] { while (run() >= 0) {} }); }
        void halt() {
            for (auto* s : ins) { s->stopReader(); }
            for (auto* s : outs) { s->stopWriter(); }
            if (worker.joinable()) { worker.join(); }
            for (auto* s : ins) { s->clearReadStop(); }
            for (auto* s : outs) { s->clearWriteStop(); }
        }
        std::vector<untyped_stream*> ins, outs;
        std::thread worker;
        bool running = false, paused = false;
        int depth = 0;
    };
}
#include "processor.h"
