// sdrplusplus_b200/csrc/design.cpp -- see design.h.  Host-only, init-time code (no CUDA).
#include "design.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <dlfcn.h>

namespace b200 {

static const double kPi = 3.14159265358979323846;   // DB_M_PI, core/src/dsp/math/constants.h:3

double hz_to_rads(double freq, double samplerate) { return 2.0 * kPi * (freq / samplerate); }

static double sinc(double x) { return (x == 0.0) ? 1.0 : (std::sin(x) / x); }          // math/sinc.h:5-7

// window::cosine (window/cosine.h:7-15)
static double cosine_window(double n, double N, const double* coefs, int count) {
    double win = 0.0, sign = 1.0;
    for (int i = 0; i < count; i++) {
        win += sign * coefs[i] * std::cos((double)i * 2.0 * kPi * n / N);
        sign = -sign;
    }
    return win;
}
static double nuttall(double n, double N) {                                             // window/nuttall.h:5-8
    const double c[] = { 0.355768, 0.487396, 0.144232, 0.012604 };
    return cosine_window(n, N, c, 4);
}
static double blackman(double n, double N) {                                            // window/blackman.h:5-8
    const double c[] = { 0.42, 0.5, 0.08 };
    return cosine_window(n, N, c, 3);
}

int estimate_tap_count(double transWidth, double samplerate) { return (int)(3.8 * samplerate / transWidth); }

// taps::windowedSinc<float> (taps/windowed_sinc.h:9-29) with window::nuttall, via taps::lowPass (low_pass.h:7-11)
std::vector<float> lowpass_taps(double cutoff, double transWidth, double samplerate, bool odd) {
    int count = estimate_tap_count(transWidth, samplerate);
    if (odd && !(count % 2)) { count++; }
    return windowed_sinc_taps(count, hz_to_rads(cutoff, samplerate));
}
// taps::windowedSinc<float>(count, omega, window::nuttall) (taps/windowed_sinc.h:9-29; the (cutoff, samplerate) overload,
// :31-34, passes omega = hzToRads(cutoff, samplerate))
std::vector<float> windowed_sinc_taps(int count, double omega) {
    std::vector<float> taps(count > 0 ? count : 0);
    const double half = (double)count / 2.0;
    const double corr = 1.0 * omega / kPi;
    for (int i = 0; i < count; i++) {
        double t = (double)i - half + 0.5;
        taps[i] = (float)(sinc(t * omega) * nuttall(t - half, count) * corr);
    }
    return taps;
}

// taps::highPass (taps/high_pass.h:7-14): sinc at fs/2 - cutoff, window nuttall(n,N) * (-1)^round(n)
std::vector<float> highpass_taps(double cutoff, double transWidth, double samplerate, bool odd) {
    int count = estimate_tap_count(transWidth, samplerate);
    if (odd && !(count % 2)) { count++; }
    std::vector<float> taps(count > 0 ? count : 0);
    const double omega = hz_to_rads((samplerate / 2.0) - cutoff, samplerate);
    const double half = (double)count / 2.0;
    const double corr = 1.0 * omega / kPi;
    for (int i = 0; i < count; i++) {
        double t = (double)i - half + 0.5;
        double n = t - half;
        double w = nuttall(n, count) * ((((int)std::round(n)) % 2) ? -1.0f : 1.0f);
        taps[i] = (float)(sinc(t * omega) * w * corr);
    }
    return taps;
}

// taps::bandPass<complex_t> (taps/band_pass.h:11-26) over windowedSinc<complex_t> (windowed_sinc.h:22-25): the window is
// phasor(-offsetOmega * n) * nuttall(n, N) with offsetOmega and n narrowed to float (complex_t * double multiplies in
// float), the sinc is narrowed to float before the complex multiply, and the final correction is a float multiply.
// Returns interleaved (re, im) pairs.
std::vector<float> bandpass_c_taps(double bandStart, double bandStop, double transWidth, double samplerate, bool odd) {
    const float offsetOmega = (float)hz_to_rads((bandStart + bandStop) / 2.0, samplerate);
    int count = estimate_tap_count(transWidth, samplerate);
    if (odd && !(count % 2)) { count++; }
    std::vector<float> taps(2 * (size_t)(count > 0 ? count : 0));
    const double omega = hz_to_rads((bandStop - bandStart) / 2.0, samplerate);
    const double half = (double)count / 2.0;
    const double corr = 1.0 * omega / kPi;
    for (int i = 0; i < count; i++) {
        const double t = (double)i - half + 0.5;
        const double n = t - half;
        const float x = -offsetOmega * (float)n;
        const float pr = cosf(x), pi = sinf(x);                    // math::phasor
        const float wn = (float)nuttall(n, count);
        const float wr = pr * wn, wi = pi * wn;
        const float cr = (float)sinc(t * omega), ci = 0.0f;
        const float re = (cr * wr) - (ci * wi), im = (ci * wr) + (cr * wi);
        taps[2 * (size_t)i] = re * (float)corr;
        taps[2 * (size_t)i + 1] = im * (float)corr;
    }
    return taps;
}

// loop::PhaseControlLoop<float>::criticallyDamped (phase_control_loop.h:27-32) with its float locals
void pll_coefficients(double bandwidth, float& alpha, float& beta) {
    const float bw = (float)bandwidth;
    const float damp = (float)(std::sqrt(2.0) / 2.0);
    const float den = (float)(1.0 + 2.0 * damp * bw + bw * bw);
    alpha = (4 * damp * bw) / den;
    beta = (4 * bw * bw) / den;
}

// clock_recovery::MM::generateInterpTaps (mm.h:168-173)
std::vector<float> mm_interp_bank(int phases, int taps) {
    const int count = phases * taps;
    const double omega = hz_to_rads(0.5 / (double)phases, 1.0);
    const double half = (double)count / 2.0;
    const double corr = (double)phases * omega / kPi;
    std::vector<float> bank((size_t)count, 0.0f);
    for (int i = 0; i < count; i++) {
        const double t = (double)i - half + 0.5;
        bank[(size_t)((phases - 1) - (i % phases)) * taps + (i / phases)] = (float)(sinc(t * omega) * nuttall(t - half, count) * corr);
    }
    return bank;
}

// iq_frontend.cpp:281-291
std::vector<float> fmif_window(int bins) {
    std::vector<float> w((size_t)bins);
    for (int i = 0; i < bins; i++) { w[i] = (float)nuttall(i, bins - 1); }
    return w;
}
std::vector<float> dft_twiddles(int n) {
    std::vector<float> t((size_t)2 * n);
    for (int k = 0; k < n; k++) {
        const double a = -2.0 * 3.14159265358979323846 * (double)k / (double)n;
        t[2 * k] = (float)std::cos(a);
        t[2 * k + 1] = (float)std::sin(a);
    }
    return t;
}
std::vector<float> fft_window(int window, int nz) {
    std::vector<float> w(nz);
    for (int i = 0; i < nz; i++) {
        float sign = (i % 2) ? -1.0f : 1.0f;
        if (window == 0) { w[i] = 1.0f * sign; }
        else if (window == 1) { w[i] = (float)(blackman(i, nz) * sign); }
        else { w[i] = (float)(nuttall(i, nz) * sign); }
    }
    return w;
}

// iq_frontend.h:59-63
void fft_frame_params(double samplerate, int size, double rate, int& nz, int& skip) {
    int interval = (int)std::round(samplerate / rate);
    nz = interval < size ? interval : size;
    skip = interval - nz;
}

// ---- decimation plan registry ----
static std::mutex g_plan_mtx;
static std::map<int, DecimPlan> g_plans;
static bool g_default_tried = false;

int register_decim_plan(int ratio, int nstages, const int* decims, const int* tapcounts, const float* const* taps) {
    if (ratio < 2 || (ratio & (ratio - 1)) || nstages < 1 || nstages > 8) { return -1; }
    DecimPlan p;
    p.ratio = ratio;
    int prod = 1;
    for (int i = 0; i < nstages; i++) {
        if (decims[i] < 1 || tapcounts[i] < 1) { return -1; }
        DecimStage st;
        st.decim = decims[i];
        st.taps.assign(taps[i], taps[i] + tapcounts[i]);
        prod *= decims[i];
        p.stages.push_back(st);
    }
    if (prod != ratio) { return -1; }
    std::lock_guard<std::mutex> lck(g_plan_mtx);
    g_plans[ratio] = p;
    return 0;
}

static std::string default_plan_path() {
    const char* env = getenv("B200_DECIM_PLANS");
    if (env) { return env; }
    Dl_info info;
    if (!dladdr((void*)&default_plan_path, &info) || !info.dli_fname) { return ""; }
    std::string p = info.dli_fname;
    size_t slash = p.rfind('/');
    if (slash == std::string::npos) { return ""; }
    // the library lives in sdrplusplus_b200/, the table in sdrplusplus_b200/data/
    return p.substr(0, slash) + "/data/decim_plans.bin";
}

int load_decim_plans(const char* path) {
    std::string p = path ? std::string(path) : default_plan_path();
    FILE* f = fopen(p.c_str(), "rb");
    if (!f) { return -1; }
    char magic[8];
    int32_t n = 0;
    int rc = -1;
    if (fread(magic, 1, 8, f) == 8 && !memcmp(magic, "SDRPPDP1", 8) && fread(&n, 4, 1, f) == 1) {
        rc = 0;
        for (int i = 0; i < n && rc == 0; i++) {
            int32_t hdr[2];
            if (fread(hdr, 4, 2, f) != 2 || hdr[1] < 1 || hdr[1] > 8) { rc = -1; break; }
            std::vector<int> D(hdr[1]), T(hdr[1]);
            std::vector<std::vector<float>> taps(hdr[1]);
            std::vector<const float*> ptrs(hdr[1]);
            for (int s = 0; s < hdr[1]; s++) {
                int32_t sh[2];
                if (fread(sh, 4, 2, f) != 2 || sh[1] < 1 || sh[1] > (1 << 20)) { rc = -1; break; }
                D[s] = sh[0];
                T[s] = sh[1];
                taps[s].resize(sh[1]);
                if (fread(taps[s].data(), 4, (size_t)sh[1], f) != (size_t)sh[1]) { rc = -1; break; }
                ptrs[s] = taps[s].data();
            }
            if (rc == 0 && register_decim_plan(hdr[0], hdr[1], D.data(), T.data(), ptrs.data())) { rc = -1; }
        }
    }
    fclose(f);
    return rc;
}

const DecimPlan* find_decim_plan(int ratio) {
    {
        std::lock_guard<std::mutex> lck(g_plan_mtx);
        auto it = g_plans.find(ratio);
        if (it != g_plans.end()) { return &it->second; }
        if (g_default_tried) { return nullptr; }
        g_default_tried = true;
    }
    load_decim_plans(nullptr);
    std::lock_guard<std::mutex> lck(g_plan_mtx);
    auto it = g_plans.find(ratio);
    return it != g_plans.end() ? &it->second : nullptr;
}

// ---- RationalResampler::reconfigure (rational_resampler.h:120-165) ----
static int igcd(int a, int b) {
    a = std::abs(a); b = std::abs(b);
    while (b) { int t = a % b; a = b; b = t; }
    return a;
}
static const int kMaxRatio = 1 << 13;   // PowerDecimator::getMaxRatio (power_decimator.h:28-30)

ResampPlan make_resamp_plan(double inSR, double outSR) {
    ResampPlan pl;
    int predecPower = (int)std::floor(std::log2(inSR / outSR));
    if (predecPower > kMaxRatio) { predecPower = kMaxRatio; }
    int predecRatio = kMaxRatio;
    if (predecPower < 0) { predecRatio = 0; }
    else if (predecPower < 31 && (1 << predecPower) < kMaxRatio) { predecRatio = 1 << predecPower; }
    double intSR = inSR;
    pl.use_decim = (inSR > outSR && predecPower > 0);
    if (pl.use_decim) { intSR = inSR / (double)predecRatio; }
    int IntSR = (int)std::round(intSR);
    int OutSR = (int)std::round(outSR);
    int g = igcd(IntSR, OutSR);
    pl.interp = OutSR / g;
    pl.decim = IntSR / g;
    pl.predec_ratio = pl.use_decim ? predecRatio : 1;
    pl.taps_per_phase = 0;
    if (pl.interp == pl.decim) {
        pl.mode = pl.use_decim ? 1 : 3;
        return pl;
    }
    double tapSamplerate = intSR * (double)pl.interp;
    double tapBandwidth = (inSR < outSR ? inSR : outSR) / 2.0;
    double tapTransWidth = tapBandwidth * 0.1;
    pl.rtaps = lowpass_taps(tapBandwidth, tapTransWidth, tapSamplerate);
    for (auto& t : pl.rtaps) { t *= (float)pl.interp; }
    pl.taps_per_phase = ((int)pl.rtaps.size() + pl.interp - 1) / pl.interp;
    pl.mode = pl.use_decim ? 0 : 2;
    return pl;
}

std::vector<float> build_polyphase_bank(int interp, const std::vector<float>& taps, int& tpp) {
    tpp = ((int)taps.size() + interp - 1) / interp;
    std::vector<float> bank((size_t)interp * tpp, 0.0f);
    int tot = interp * tpp;
    for (int i = 0; i < tot; i++) {
        bank[(size_t)((interp - 1) - (i % interp)) * tpp + (i / interp)] = (i < (int)taps.size()) ? taps[i] : 0.0f;
    }
    return bank;
}

}
