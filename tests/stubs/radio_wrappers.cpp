// tests/stubs/radio_wrappers.cpp -- TEST INFRASTRUCTURE.  Compiles the reference's OWN radio-module demodulator wrappers
// (decoder_modules/radio/src/demodulators/{nfm,am,usb,lsb,dsb,raw}.h, read where they lie through a symlink farm that
// tests/test_boundary_compile.py builds) against sdrplusplus_b200/host/dsp, unchanged, and drives each through the
// demod::Demodulator interface the radio module uses (radio_module.h:419-562): init -> start -> chunks in -> audio out.
// With a CUDA device it prints "<name> <samples> <sum |l|+|r|>" per demodulator; without one it reports that every
// block refused to compute (no CPU fallback) and exits 0: the compile and link are the CPU-side check.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include "demod.h"

template <class D>
static int drive(const char* name, const std::vector<dsp::complex_t>& iq, int chunk, double bw) {
    ConfigManager cfg;
    dsp::stream<dsp::complex_t> in(chunk);
    D d;
    d.init("Radio", &cfg, &in, bw, 48000.0);
    d.setBandwidth(bw);
    d.start();
    std::thread writer([&] {
        for (size_t i = 0; i + chunk <= iq.size(); i += (size_t)chunk) {
            memcpy(in.writeBuf, &iq[i], (size_t)chunk * sizeof(dsp::complex_t));
            if (!in.swap(chunk)) { return; }
        }
    });
    double cs = 0.0;
    size_t total = 0;
    dsp::stream<dsp::stereo_t>* out = d.getOutput();
    for (size_t c = 0; c < iq.size() / (size_t)chunk; c++) {
        int n = out->read();
        if (n < 0) { break; }
        for (int i = 0; i < n; i++) { cs += std::fabs(out->readBuf[i].l) + std::fabs(out->readBuf[i].r); }
        total += (size_t)n;
        out->flush();
    }
    writer.join();
    d.stop();
    printf("%s %zu %.9e if=%.0f\n", name, total, cs, d.getIFSampleRate());
    return 0;
}

int main(int argc, char** argv) {
    const bool have_gpu = b200_device_count() > 0 && b200_init(0) == 0;
    if (!have_gpu) {
        // no device: creation fails loudly, nothing computes (there is no CPU fallback)
        ConfigManager cfg;
        dsp::stream<dsp::complex_t> in(1000);
        demod::NFM d;
        d.init("Radio", &cfg, &in, 12500.0, 48000.0);
        printf("no CUDA device: %s\n", b200_last_error());
        return 0;
    }
    if (argc < 2) { fprintf(stderr, "usage: %s iq.f32\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { return 2; }
    fseek(f, 0, SEEK_END);
    long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<dsp::complex_t> iq((size_t)bytes / sizeof(dsp::complex_t));
    if (fread(iq.data(), sizeof(dsp::complex_t), iq.size(), f) != iq.size()) { return 2; }
    fclose(f);
    // the wrappers sit BEHIND the VFO: their input is already at the demodulator's IF rate
    drive<demod::NFM>("NFM", iq, 5000, 12500.0);
    drive<demod::AM>("AM", iq, 5000, 10000.0);
    drive<demod::USB>("USB", iq, 5000, 2800.0);
    drive<demod::LSB>("LSB", iq, 5000, 2800.0);
    drive<demod::DSB>("DSB", iq, 5000, 4600.0);
    drive<demod::RAW>("RAW", iq, 5000, 48000.0);
    return 0;
}
