// host/dsp/b200/handle.h -- what every per-block adapter shares: one libb200dsp stand-alone block handle
// (b200_*_create / b200_block_process / b200_block_reset / b200_block_destroy, include/b200dsp.h).  Parameters the C ABI
// fixes at creation (rates, modes, coefficient sets) are changed by building a new handle under the block's control
// mutex with the worker paused -- the point at which the reference's own setters run (tempStop ... tempStart).
#pragma once
#include "../../../../include/b200dsp.h"

namespace dsp::b200 {
    class Handle {
    public:
        Handle() {}
        Handle(const Handle&) = delete;
        Handle& operator=(const Handle&) = delete;
        ~Handle() { b200_block_destroy(h); }
        // replaces the current block by `fresh` (nullptr = creation failed: the old block is dropped, ok() turns false)
        void adopt(b200_block* fresh) {
            b200_block_destroy(h);
            h = fresh;
        }
        bool ok() const { return h != nullptr; }
        b200_block* get() const { return h; }
        int process(int count, const void* in, void* out) { return h ? b200_block_process(h, count, in, out) : B200_ESTATE; }
        void reset() { if (h) { b200_block_reset(h); } }
    private:
        b200_block* h = nullptr;
    };
}
