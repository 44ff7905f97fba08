// tests/stubs/radio_wfm.cpp -- TEST INFRASTRUCTURE.  The reference's OWN WFM demodulator wrapper of the radio module
// (decoder_modules/radio/src/demodulators/wfm.h, read where it lies through the symlink farm tests/test_boundary_compile.py
// builds), compiled unchanged against sdrplusplus_b200/host: dsp::demod::BroadcastFM with its rdsOut stream, RDSDemod
// (host/radio/rds_demod.h in the place of the module's own file), dsp::sink::Handler, dsp::buffer::Reshaper -- and linked with
// the reference's own RDS group decoder (rds.cpp, compiled from where it lies).  Driven the way radio_module.h drives a
// demodulator: init -> start -> IF chunks in -> audio out; the decoded programme service name is read back the way the
// module shows it, through the waterfall's FFT-redraw event ("RDS: <name>" drawn into the window's draw list).
// With a CUDA device it prints "WFM <audio samples> <sum |l|+|r|>" and "RDS <text>"; without one it reports
// that the blocks refused to compute (no CPU fallback) and exits 0: compile and link are the CPU-side check.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <utils/optionlist.h>
#include <gui/widgets/waterfall.h>
#include <dsp/sink/handler_sink.h>
#include <dsp/buffer/reshaper.h>
#include "demod.h"

int main(int argc, char** argv) {
    const bool have_gpu = b200_device_count() > 0 && b200_init(0) == 0;
    ConfigManager cfg;
    cfg.conf["Radio"]["WFM"]["rds"] = true;
    cfg.conf["Radio"]["WFM"]["rdsInfo"] = true;            // soft symbols on: the Reshaper / symbol-display branch runs too
    cfg.conf["Radio"]["WFM"]["stereo"] = false;
    const int chunk = 12500;
    dsp::stream<dsp::complex_t> in(chunk);
    if (!have_gpu) {
        demod::WFM d;
        d.init("Radio", &cfg, &in, 150000.0, 48000.0);
        printf("no CUDA device: %s\n", b200_last_error());
        return 0;
    }
    if (argc < 2) { fprintf(stderr, "usage: %s if_iq.f32\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { return 2; }
    fseek(f, 0, SEEK_END);
    long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<dsp::complex_t> iq((size_t)bytes / sizeof(dsp::complex_t));
    if (fread(iq.data(), sizeof(dsp::complex_t), iq.size(), f) != iq.size()) { return 2; }
    fclose(f);

    demod::WFM d;
    d.init("Radio", &cfg, &in, 150000.0, 48000.0);
    d.start();
    const size_t nch = iq.size() / (size_t)chunk;
    std::thread writer([&] {
        for (size_t i = 0; i < nch; i++) {
            memcpy(in.writeBuf, &iq[i * (size_t)chunk], (size_t)chunk * sizeof(dsp::complex_t));
            if (!in.swap(chunk)) { return; }
        }
    });
    double cs = 0.0;
    size_t total = 0;
    dsp::stream<dsp::stereo_t>* out = d.getOutput();
    for (size_t c = 0; c < nch; c++) {
        int n = out->read();
        if (n < 0) { break; }
        for (int i = 0; i < n; i++) { cs += std::fabs(out->readBuf[i].l) + std::fabs(out->readBuf[i].r); }
        total += (size_t)n;
        out->flush();
    }
    writer.join();
    // let the RDS branch (three more worker threads behind rdsOut) drain what is in flight, then read the name back the way
    // the module displays it
    std::this_thread::sleep_for(std::chrono::milliseconds(300));
    ImGuiWindow win;
    ImGui::WaterFall::FFTRedrawArgs args;
    args.window = &win;
    gui::waterfall.onFFTRedraw.emit(args);
    d.stop();
    printf("WFM %zu %.9e\n", total, cs);
    printf("RDS %s\n", win.list.texts ? win.list.last_text.c_str() : "(nothing decoded)");
    return 0;
}
