// cubicsdr_b200/csrc/design.h -- host-side filter design and resampler planning for the B200 SDR engine.
//
// Everything the reference obtains from liquid-dsp 1.5.0 `*_create` calls on the hot path is designed here,
// from the published algorithms (Kaiser-windowed sinc, Butterworth/bilinear, half-band polyphase split) and
// the structure documented in SURVEY.md Appendix A.  No liquid code is used; citations give the reference
// call site each table serves.  Tables are built once per (rate, bandwidth) plan and uploaded to HBM.
#pragma once
#include <cstdint>
#include <vector>

namespace sdrb200 {

// ---- primitives ------------------------------------------------------------------------------------------
float kaiser_beta_for_atten(float as_db);                       // Vaidyanathan fit, used by every Kaiser design below
unsigned filter_len_estimate(float df, float as_db);            // (As-7.95)/(14.26 df), truncated
void kaiser_lowpass(unsigned n, float fc, float as_db, float mu, float *h);   // sinc(2 fc t) * kaiser, centre tap 1
// firpfbch_crcf_create_kaiser(LIQUID_ANALYZER, M, m, As): branch filters of the analysis channelizer and its DFT twiddles.
// hq[c * 2m + q] = h[(M - 1 - c) + q M] with h = kaiser_lowpass(2 M m + 1, 0.5 / M, As, 0);  tw[n] = exp(-j 2 pi n / M)
void pfb_analysis_tables(unsigned M, unsigned m, float as_db, std::vector<float> &hq, std::vector<float> &tw_re_im);
// firpfbch2_crcf_create_kaiser(LIQUID_ANALYZER, M, m, As): hq[cm * 4m + d] = g[cm + (M/2) d], g = h / sum(h),
// h = kaiser_lowpass(2 M m + 1, 1 / M, As, 0);  tw[n] = exp(+j 2 pi n / M) (the analyzer ends in an inverse DFT scaled by 1/M)
void pfb2_analysis_tables(unsigned M, unsigned m, float as_db, std::vector<float> &hq, std::vector<float> &tw_re_im);
// SDRThread::getOptimalChannelCount (src/sdr/SoapySDRThread.cpp:676-693), CHANNELIZER_RATE_MAX = 500000
int optimal_channel_count(long long sample_rate);

// 1024-entry sine table of the table-driven oscillator (nco_crcf: DemodulatorPreThread.cpp:22, ModemUSB.cpp:9)
void nco_sine_table(float *tab1024);
// frequency word: what nco_crcf_set_frequency(float) stores.  `dtheta_rad` is already narrowed to float by the caller.
uint32_t nco_phase_step(float dtheta_rad);
// the value DemodulatorPreThread.cpp:157 passes: 2*pi*|shift|/fs evaluated in double, narrowed to float
float nco_shift_word_arg(long long shift_hz, long long sample_rate);

// ---- half-band stages (msresamp2 inside msresamp_crcf / msresamp_rrrf) ---------------------------------------
// Polyphase "filter branch" taps h1[0..2m) of the half-band used at cascade position g (g = 0 is the LAST stage
// executed when decimating, the FIRST when interpolating).  Only m in {3,5,10} occurs at As = 60 dB.
struct HalfbandPlan {
    int num_stages = 0;                  // S
    std::vector<int> m;                  // m[g], g = 0..S-1
};
HalfbandPlan halfband_plan(int num_stages, float as_db);
// h1 branch taps (length 2m) for semi-length m at the stop-band the cascade uses (As+5).  Returns false if the
// (m, As) pair is not one of the Parks-McClellan prototypes this engine carries.
bool halfband_branch_taps(int m, float as_db, float *h1);

// ---- multi-stage resampler plan (msresamp_crcf_create(r, 60): DemodulatorWorkerThread.cpp:100;
//      msresamp_rrrf_create(r, 60): ModemAnalog.cpp:30) ------------------------------------------------------
struct ResamplerPlan {
    bool interp = false;                 // r > 1
    float rate = 1.f;                    // r as float
    int num_halfband = 0;                // S
    float rate_arb = 1.f;                // r * 2^S (decim) or r / 2^S (interp)
    float zeta = 1.f;                    // 1 / 2^S, applied to decimator output only
    HalfbandPlan hb;
    uint32_t arb_step = 1u << 24;        // fixed-point output stride, 2^24 / rate_arb
    unsigned arb_m = 7, arb_npfb = 256;
    float arb_fc = 0.f;
    std::vector<float> pfb;              // [npfb][2*arb_m] taps, pfb[a*14 + n] multiplies x[t-n]
    float delay() const;                 // msresamp_get_delay equivalent (informational)
};
ResamplerPlan make_resampler_plan(float rate, float as_db);
// closed forms for the arbitrary stage: number of outputs produced by `n_in` inputs starting from phase `phase`
// (phase as liquid keeps it between calls), and the phase afterwards
uint64_t arb_output_count(uint32_t step, uint64_t phase, uint64_t n_in, uint64_t *phase_after);

// ---- demodulator filters ---------------------------------------------------------------------------------------
void am_dc_block_taps(unsigned m, float as_db, float *h /*2m+1*/);          // firfilt_rrrf_create_dc_blocker(25,30): ModemAM.cpp:9
void hilbert_quadrature_taps(unsigned m, float as_db, float *hq /*2m*/);    // firhilbf_create(5,90): ModemUSB.cpp:11
struct Sos { float b[3]; float a[3]; };
void butterworth_lowpass_sos(unsigned order, float fc, std::vector<Sos> &sos);   // iirfilt_crcf_create_lowpass(6,0.25): ModemUSB.cpp:8
// iirfilt_crcf_create_prototype(LIQUID_IIRDES_CHEBY2, LIQUID_IIRDES_BANDPASS, LIQUID_IIRDES_SOS, 5, fc, f0, 1, 60): the 19 kHz pilot
// band-pass of ModemFMStereo.cpp:130-136 as five second-order sections.  The poles sit 0.002 inside the unit circle, so the sections
// must be the reference's to the last bit: the arithmetic below follows liquid 1.5.0's iirdes stage by stage (analog prototype as
// constants, bilinear_zpkf in double, iirdes_dzpk_lp2bp and iirdes_dzpk2sosf in float) and is pinned against the binary.
void cheby2_bandpass5_sos(float fc, float f0, std::vector<Sos> &sos);
void sos_section_object(Sos &s);   // the coefficients as iirfiltsos_crcf_create stores them (see design.cpp)
// ModemFMStereo.cpp:149-159: first-order de-emphasis section {b0, b1, a1} for a time constant in microseconds
void fm_deemphasis(int tau_us, int audio_rate, float *b2, float *a1);

}  // namespace sdrb200
