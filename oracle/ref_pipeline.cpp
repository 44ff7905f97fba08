/*
 * oracle/ref_pipeline.cpp  --  TEST/BENCH INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The timed CPU baseline: the reference's OWN blocks (headers under /root/reference/core/src, included with -I,
 * nothing copied), wired exactly like IQFrontEnd + the radio module's WFM chain and driven the way the
 * reference's own (never instantiated) dsp::bench::SpeedTester drives a graph (speed_tester.h:31-56,78-84):
 * a pre-filled uniform[-1,1) buffer is swap()ed into the graph as fast as it is accepted for a fixed wall
 * time; throughput = samples accepted / time.  One std::thread per block, as in the reference
 * (core/src/dsp/block.h:71-73).
 *
 *   writer -> Splitter -> { Reshaper(keep nz, skip) -> Handler: window*(-1)^n, FFT, 10log10(|X/N|^2) }
 *                      -> N x { RxVFO -> BroadcastFM (mono, low-pass) -> Null sink }
 *
 * VOLK and FFTW are absent from this image: the leaf kernels are the scalar restatement of oracle/shim
 * compiled -O3 -march=native (the compiler vectorises what VOLK would dispatch by hand).  State that next to
 * every number this program prints.  The FFT handler restates IQFrontEnd::handler (iq_frontend.cpp:248-267).
 *
 * usage: ref_pipeline <samplerate> <chunk> <fft_size> <fft_rate> <n_vfo> <duration_ms> [mode]
 *        mode: "wfm" (default) | "fft" (spectrum branch only) | "vfo" (VFO branch only)
 * prints one line: samples_per_second threads
 */
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>
#include <unistd.h>
#include <vector>

#include <dsp/types.h>
#include <dsp/routing/splitter.h>
#include <dsp/buffer/reshaper.h>
#include <dsp/sink/handler_sink.h>
#include <dsp/sink/null_sink.h>
#include <dsp/channel/rx_vfo.h>
#include <dsp/demod/broadcast_fm.h>
#include <dsp/window/nuttall.h>
#include <fftw3.h>

using namespace dsp;

struct FftCtx {
    int size, nz;
    float* window;
    fftwf_complex* in;
    fftwf_complex* out;
    fftwf_plan plan;
    float* line;
};

// IQFrontEnd::handler (core/src/signal_path/iq_frontend.cpp:248-267)
static void fftHandler(complex_t* data, int count, void* ctx) {
    FftCtx* f = (FftCtx*)ctx;
    volk_32fc_32f_multiply_32fc((lv_32fc_t*)f->in, (lv_32fc_t*)data, f->window, f->nz);
    fftwf_execute(f->plan);
    volk_32fc_s32f_power_spectrum_32f(f->line, (lv_32fc_t*)f->out, f->size, f->size);
}

int main(int argc, char** argv) {
    if (argc < 7) {
        fprintf(stderr, "usage: %s samplerate chunk fft_size fft_rate n_vfo duration_ms [wfm|fft|vfo]\n", argv[0]);
        return 2;
    }
    const double fs = atof(argv[1]);
    const int chunk = atoi(argv[2]);
    const int fftSize = atoi(argv[3]);
    const double fftRate = atof(argv[4]);
    const int nvfo = atoi(argv[5]);
    const int durationMs = atoi(argv[6]);
    const char* mode = argc > 7 ? argv[7] : "wfm";
    const bool doFft = strcmp(mode, "vfo") != 0 && fftSize > 0;
    const bool doVfo = strcmp(mode, "fft") != 0 && nvfo > 0;
    if (chunk > STREAM_BUFFER_SIZE) { fprintf(stderr, "chunk exceeds STREAM_BUFFER_SIZE\n"); return 2; }

    // input: the SpeedTester distribution, seeded
    stream<complex_t> input;
    std::vector<complex_t> buf(chunk);
    std::mt19937 rng(0x5D12);
    std::uniform_real_distribution<float> dist(-1.0f, 1.0f);
    for (auto& s : buf) { s.re = dist(rng); s.im = dist(rng); }

    routing::Splitter<complex_t> split;
    split.init(&input);
    int threads = 2;   // writer + splitter

    // spectrum branch (IQFrontEnd::init / updateFFTPath, iq_frontend.cpp:59-72,269-309)
    stream<complex_t> fftIn;
    buffer::Reshaper<complex_t> reshape;
    sink::Handler<complex_t> fftSink;
    FftCtx f;
    if (doFft) {
        int fftInterval = round(fs / fftRate);
        f.size = fftSize;
        f.nz = std::min<int>(fftInterval, fftSize);
        int skip = fftInterval - f.nz;
        f.window = buffer::alloc<float>(f.nz);
        for (int i = 0; i < f.nz; i++) { f.window[i] = window::nuttall(i, f.nz) * ((i % 2) ? -1.0f : 1.0f); }
        f.in = (fftwf_complex*)fftwf_malloc(fftSize * sizeof(fftwf_complex));
        f.out = (fftwf_complex*)fftwf_malloc(fftSize * sizeof(fftwf_complex));
        f.plan = fftwf_plan_dft_1d(fftSize, f.in, f.out, FFTW_FORWARD, FFTW_ESTIMATE);
        buffer::clear(f.in, fftSize - f.nz, f.nz);
        f.line = buffer::alloc<float>(fftSize);
        if (f.nz > STREAM_BUFFER_SIZE) { reshape.out.setBufferSize(f.nz); }
        reshape.init(&fftIn, f.nz, skip);
        fftSink.init(&reshape.out, fftHandler, &f);
        split.bindStream(&fftIn);
        threads += 3;   // reshaper (2 threads) + handler
    }

    // VFO branch: RxVFO -> BroadcastFM(mono, lowpass) -> Null   (radio_module.h:80-125, wfm.h:78,363-365)
    std::vector<stream<complex_t>*> vfoIn;
    std::vector<channel::RxVFO*> vfos;
    std::vector<demod::BroadcastFM*> demods;
    std::vector<sink::Null<stereo_t>*> sinks;
    if (doVfo) {
        for (int v = 0; v < nvfo; v++) {
            // offsets on a grid inside +-0.4 fs, like BASELINE config 2 (+-5, +-15, +-25, +-35 MHz at 100 MS/s)
            double off = ((v / 2) * 2 + 1) * (fs * 0.05) * ((v % 2) ? -1.0 : 1.0);
            auto* in = new stream<complex_t>;
            auto* vfo = new channel::RxVFO(in, fs, 250000.0, 150000.0, off);
            auto* dm = new demod::BroadcastFM(&vfo->out, 75000.0, 250000.0, false, true, false);
            auto* ns = new sink::Null<stereo_t>;
            ns->init(&dm->out);
            split.bindStream(in);
            vfoIn.push_back(in); vfos.push_back(vfo); demods.push_back(dm); sinks.push_back(ns);
            threads += 3;
        }
    }

    // start consumers first, like IQFrontEnd::start
    for (auto* s : sinks) { s->start(); }
    for (auto* d : demods) { d->start(); }
    for (auto* v : vfos) { v->start(); }
    if (doFft) { fftSink.start(); reshape.start(); }
    split.start();

    std::atomic<bool> stop{ false };
    std::atomic<long long> accepted{ 0 };
    std::thread writer([&] {
        while (!stop.load(std::memory_order_relaxed)) {
            memcpy(input.writeBuf, buf.data(), (size_t)chunk * sizeof(complex_t));
            if (!input.swap(chunk)) { break; }
            accepted.fetch_add(chunk, std::memory_order_relaxed);
        }
    });

    // warm up for 20% of the run, then count
    std::this_thread::sleep_for(std::chrono::milliseconds(durationMs / 5));
    long long a0 = accepted.load();
    auto t0 = std::chrono::steady_clock::now();
    std::this_thread::sleep_for(std::chrono::milliseconds(durationMs));
    long long a1 = accepted.load();
    auto t1 = std::chrono::steady_clock::now();
    double secs = std::chrono::duration<double>(t1 - t0).count();

    stop = true;
    input.stopWriter();
    writer.join();
    split.stop();
    if (doFft) { reshape.stop(); fftSink.stop(); }
    for (auto* v : vfos) { v->stop(); }
    for (auto* d : demods) { d->stop(); }
    for (auto* s : sinks) { s->stop(); }

    printf("%.6e %d\n", (double)(a1 - a0) / secs, threads);
    fflush(stdout);
    _exit(0);   // the reference's block destructors join threads in an order that can hang; we are done
}
