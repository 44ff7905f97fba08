// sdrplusplus_b200/csrc/design.h -- host-side filter/window/plan design (fp64 -> fp32), the same formulas
// the reference evaluates at init time.  Compiled with -ffp-contract=off so the
// results are bit-identical to the reference's (checked against the oracle in tests/test_design.py).
#pragma once
#include <vector>
#include <cstdint>

namespace b200 {

double hz_to_rads(double freq, double samplerate);                       // math::hzToRads (hz_to_rads.h:6-8)
int estimate_tap_count(double transWidth, double samplerate);             // taps::estimateTapCount
std::vector<float> lowpass_taps(double cutoff, double transWidth, double samplerate, bool odd = false); // taps::lowPass
std::vector<float> windowed_sinc_taps(int count, double omega);          // taps::windowedSinc<float> with window::nuttall
std::vector<float> highpass_taps(double cutoff, double transWidth, double samplerate, bool odd = false);                 // taps::highPass
std::vector<float> bandpass_c_taps(double bandStart, double bandStop, double transWidth, double samplerate, bool odd = false); // taps::bandPass<complex_t>, (re, im) pairs
void pll_coefficients(double bandwidth, float& alpha, float& beta);
// clock_recovery::MM::generateInterpTaps (mm.h:168-173): windowedSinc<float>(phases * taps, hzToRads(0.5 / phases, 1), nuttall,
// norm = phases) laid out by buildPolyphaseBank (polyphase_bank.h:15-48) as [phases][taps]
std::vector<float> mm_interp_bank(int phases, int taps);
std::vector<float> fmif_window(int bins);                                // noise_reduction::FMIF::initBuffers: window::nuttall(i, bins - 1) (fm_if.h:116)
std::vector<float> dft_twiddles(int n);                                  // exp(-2 pi i k / n) as (re, im) pairs, fp64 -> fp32      // PhaseControlLoop<float>::criticallyDamped
std::vector<float> fft_window(int window, int nz);                       // IQFrontEnd::updateFFTPath window * (-1)^i
void fft_frame_params(double samplerate, int size, double rate, int& nz, int& skip); // genReshapeParams

struct DecimStage { int decim; std::vector<float> taps; };
struct DecimPlan { int ratio; std::vector<DecimStage> stages; };
// registry of the reference's power-of-two decimation plans (decim/plans.h)
int register_decim_plan(int ratio, int nstages, const int* decims, const int* tapcounts, const float* const* taps);
int load_decim_plans(const char* path);          // flat table written by tools/extract_decim_plans.py
const DecimPlan* find_decim_plan(int ratio);     // lazy default load; nullptr when unavailable

struct ResampPlan {
    int mode;            // 0 BOTH, 1 DECIM_ONLY, 2 RESAMP_ONLY, 3 NONE
    bool use_decim;
    int predec_ratio;
    int interp, decim;
    std::vector<float> rtaps;   // prototype * interp (empty when the polyphase stage is bypassed)
    int taps_per_phase;
};
ResampPlan make_resamp_plan(double inSR, double outSR);   // RationalResampler::reconfigure
// polyphase bank [interp][tapsPerPhase]  (polyphase_bank.h:15-48: phase-reversed fill)
std::vector<float> build_polyphase_bank(int interp, const std::vector<float>& taps, int& tapsPerPhase);

}
