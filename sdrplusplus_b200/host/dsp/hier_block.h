// host/dsp/hier_block.h -- dsp::hier_block: a group of blocks started and stopped together
// (core/src/dsp/hier_block.h:5-78).  Members are kept in registration order; temporary stops nest.
#pragma once
#include "block.h"

namespace dsp {
    class hier_block : public generic_block {
    public:
        virtual void init() {}
        ~hier_block() override { if (_block_init) { stop(); } }
        void start() override {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            if (!live) { live = true; each(true); }
        }
        void stop() override {
            std::lock_guard<std::recursive_mutex> lk(ctrlMtx);
            if (live) { each(false); live = false; }
        }
        void tempStop() {
            if (holds++ == 0 && live) { each(false); held = true; }
        }
        void tempStart() {
            if (holds > 0 && --holds == 0 && held) { each(true); held = false; }
        }

    protected:
        void registerBlock(generic_block* b) { members.push_back(b); }
        void unregisterBlock(generic_block* b) { members.erase(std::remove(members.begin(), members.end(), b), members.end()); }
        bool _block_init = false;
        std::recursive_mutex ctrlMtx;

    private:
        void each(bool go) {
            for (generic_block* b : members) { go ? b->start() : b->stop(); }
        }
        std::vector<generic_block*> members;
        bool live = false, held = false;
        int holds = 0;
    };
}
