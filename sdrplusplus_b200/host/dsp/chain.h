// host/dsp/chain.h -- dsp::chain<T>: an ordered list of same-type processors, each switchable at run time; the chain's
// `out` follows the last enabled block (core/src/dsp/chain.h:8-194; the radio module's IF / AF chains,
// decoder_modules/radio/src/radio_module.h:88-110).  Written around one rewire() that walks the enabled entries.
#pragma once
#include <stdexcept>
#include <vector>
#include "processor.h"

namespace dsp {
    template <class T>
    class chain {
        using proc = Processor<T, T>;
        struct entry { proc* p; bool on; stream<T>* wired; };
    public:
        chain() {}
        explicit chain(stream<T>* in) { init(in); }
        void init(stream<T>* in) { src = in; out = in; }

        template <typename Func>
        void setInput(stream<T>* in, Func onOutputChange) {
            src = in;
            rewire(onOutputChange);
        }
        void addBlock(proc* b, bool enabled) {
            if (find(b) >= 0) { throw std::runtime_error("[chain] Tried to add a block that is already part of the chain"); }
            list.push_back({ b, false, nullptr });
            if (enabled) { enableBlock(b, [](stream<T>*) {}); }
        }
        template <typename Func>
        void removeBlock(proc* b, Func onOutputChange) {
            const int i = need(b, "remove");
            disableBlock(b, onOutputChange);
            list.erase(list.begin() + i);
        }
        template <typename Func>
        void enableBlock(proc* b, Func onOutputChange) {
            entry& e = list[need(b, "enable")];
            if (e.on) { return; }
            e.on = true;
            rewire(onOutputChange);
            if (live) { b->start(); }
        }
        template <typename Func>
        void disableBlock(proc* b, Func onOutputChange) {
            entry& e = list[need(b, "disable")];
            if (!e.on) { return; }
            b->stop();
            e.on = false;
            rewire(onOutputChange);
        }
        template <typename Func>
        void setBlockEnabled(proc* b, bool enabled, Func onOutputChange) {
            if (enabled) { enableBlock(b, onOutputChange); } else { disableBlock(b, onOutputChange); }
        }
        template <typename Func>
        void enableAllBlocks(Func onOutputChange) {
            for (size_t i = 0; i < list.size(); i++) { enableBlock(list[i].p, onOutputChange); }
        }
        template <typename Func>
        void disableAllBlocks(Func onOutputChange) {
            for (size_t i = 0; i < list.size(); i++) { disableBlock(list[i].p, onOutputChange); }
        }
        void start() {
            if (live) { return; }
            for (entry& e : list) { if (e.on) { e.p->start(); } }
            live = true;
        }
        void stop() {
            if (!live) { return; }
            for (entry& e : list) { if (e.on) { e.p->stop(); } }
            live = false;
        }
        stream<T>* out = nullptr;

    private:
        int find(proc* b) const {
            for (size_t i = 0; i < list.size(); i++) { if (list[i].p == b) { return (int)i; } }
            return -1;
        }
        int need(proc* b, const char* what) const {
            const int i = find(b);
            if (i < 0) { throw std::runtime_error(std::string("[chain] Tried to ") + what + " a block that isn't part of the chain"); }
            return i;
        }
        // every enabled block reads the previous enabled block (or the chain input); report a changed tail
        template <typename Func>
        void rewire(Func onOutputChange) {
            stream<T>* cur = src;
            for (entry& e : list) {
                if (!e.on) { continue; }
                if (e.wired != cur) { e.p->setInput(cur); e.wired = cur; }
                cur = &e.p->out;
            }
            if (cur != out) { out = cur; onOutputChange(out); }
        }
        stream<T>* src = nullptr;
        std::vector<entry> list;
        bool live = false;
    };
}
