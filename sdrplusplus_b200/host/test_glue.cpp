// host/test_glue.cpp -- the two glue blocks of the operator API that need no device (dsp::buffer::Reshaper, dsp::sink::Handler)
// run as worker-thread blocks on the CPU: a writer swaps chunks of a ramp into a stream, Reshaper re-frames it, Handler hands
// every block to a callback.  usage: test_glue <keep> <skip> <total> <chunk>  ->  prints the blocks, one per line.
// tests/test_host_adapter.py compares them with the framing rule of core/src/dsp/buffer/reshaper.h.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#include "dsp/buffer/reshaper.h"
#include "dsp/sink/handler_sink.h"

// this program is NOT linked with libb200dsp: the streams' pinned buffers come from plain malloc here, so that the glue blocks
// can be exercised on a machine without a device (nothing in it computes)
extern "C" void* b200_host_alloc(uint64_t bytes) { return malloc(bytes ? bytes : 16); }
extern "C" void b200_host_free(void* p) { free(p); }

static std::mutex g_mtx;
static std::vector<std::vector<float>> g_blocks;
static void on_block(float* data, int count, void*) {
    std::lock_guard<std::mutex> lk(g_mtx);
    g_blocks.emplace_back(data, data + count);
}

int main(int argc, char** argv) {
    if (argc < 5) { return 2; }
    const int keep = atoi(argv[1]), skip = atoi(argv[2]), total = atoi(argv[3]), chunk = atoi(argv[4]);
    dsp::stream<float> in(chunk);
    dsp::buffer::Reshaper<float> rs(&in, keep, skip);
    dsp::sink::Handler<float> hs(&rs.out, on_block, nullptr);
    hs.start();
    rs.start();
    for (int i = 0; i < total; i += chunk) {
        const int n = std::min(chunk, total - i);
        for (int k = 0; k < n; k++) { in.writeBuf[k] = (float)(i + k + 1); }
        if (!in.swap(n)) { return 3; }
    }
    // how many blocks the rule yields for `total` samples
    const int carry = skip < 0 ? std::min(-skip, keep) : 0, fresh = keep - carry, gap = skip > 0 ? skip : 0;
    long long expect = 0;
    for (long long used = 0; used + fresh <= total; used += fresh + gap) { expect++; }
    for (int spin = 0; spin < 2000; spin++) {
        { std::lock_guard<std::mutex> lk(g_mtx); if ((long long)g_blocks.size() >= expect) { break; } }
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }
    rs.stop();
    hs.stop();
    for (auto& b : g_blocks) {
        for (size_t i = 0; i < b.size(); i++) { printf(i ? " %.0f" : "%.0f", b[i]); }
        printf("\n");
    }
    return 0;
}
