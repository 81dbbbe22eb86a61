// cubicsdr_b200/csrc/design.cpp -- see design.h.  Host-only, plain C++14.
#include "design.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <xmmintrin.h>
#include <cstring>

namespace sdrb200 {

static const double kPi = 3.14159265358979323846;

// ---- Kaiser machinery --------------------------------------------------------------------------------------
float kaiser_beta_for_atten(float as_db) {
    as_db = std::fabs(as_db);
    if (as_db > 50.0f) return 0.1102f * (as_db - 8.7f);
    if (as_db > 21.0f) return 0.5842f * std::pow(as_db - 21.0f, 0.4f) + 0.07886f * (as_db - 21.0f);
    return 0.0f;
}

unsigned filter_len_estimate(float df, float as_db) {
    return (unsigned)((as_db - 7.95f) / (14.26f * df));
}

// I0(z) by its power series, in double (the reference evaluates a 32-term log-gamma form in float; the two agree
// to float rounding, which is what the 1e-6 table tolerance in tests/test_design.py checks)
static double bessel_i0(double z) {
    double q = 0.25 * z * z, term = 1.0, sum = 1.0;
    for (int k = 1; k < 64; ++k) {
        term *= q / ((double)k * (double)k);
        sum += term;
        if (term < 1e-20 * sum) break;
    }
    return sum;
}

static float kaiser_window(unsigned i, unsigned n, float beta) {
    float t = (float)i - (float)(n - 1) / 2;
    float r = 2.0f * t / (float)(n - 1);
    double a = bessel_i0((double)beta * std::sqrt(std::max(0.0, 1.0 - (double)r * (double)r)));
    double b = bessel_i0((double)beta);
    return (float)(a / b);
}

static float sinc_pi(float x) {
    if (std::fabs(x) < 0.01f)
        return (float)(std::cos(kPi * x / 2.0) * std::cos(kPi * x / 4.0) * std::cos(kPi * x / 8.0));
    return (float)(std::sin(kPi * (double)x) / (kPi * (double)x));
}

void kaiser_lowpass(unsigned n, float fc, float as_db, float mu, float *h) {
    float beta = kaiser_beta_for_atten(as_db);
    for (unsigned i = 0; i < n; ++i) {
        float t = (float)i - (float)(n - 1) / 2 + mu;
        h[i] = sinc_pi(2.0f * fc * t) * kaiser_window(i, n, beta);
    }
}

// ---- table oscillator ----------------------------------------------------------------------------------------
void nco_sine_table(float *tab) {
    // sintab[i] = sinf((float)i * (2 pi / 1024 as float)); libm sinf on the host so that the table is
    // bit-identical to the one the reference builds (device sinf would differ in the last bit)
    const float dphi = 0.0061359233f;
    for (int i = 0; i < 1024; ++i) tab[i] = sinf((float)i * dphi);
}

uint32_t nco_phase_step(float dtheta_rad) {
    // SURVEY.md A.1 set_frequency: scale by 1/(2 pi), keep the fractional turn in [0,1), 32-bit phase word
    float x = (float)((double)dtheta_rad * 0.159154943091895);
    x -= truncf(x);
    if (x < 0.f) x += 1.0f;
    return (uint32_t)(int64_t)(x * 4294967296.0f);
}

float nco_shift_word_arg(long long shift_hz, long long sample_rate) {
    long long a = shift_hz < 0 ? -shift_hz : shift_hz;
    return (float)((2.0 * kPi) * (((double)a) / ((double)sample_rate)));   // DemodulatorPreThread.cpp:157
}

// ---- half-bands ------------------------------------------------------------------------------------------------
HalfbandPlan halfband_plan(int num_stages, float as_db) {
    // per-stage transition widths from fc0 = 0.4, lengths from the Kaiser estimate at As+5
    HalfbandPlan p;
    p.num_stages = num_stages;
    float fc = 0.4f, as_s = as_db + 5.0f;
    for (int i = 0; i < num_stages; ++i) {
        fc = (i == 1) ? (0.5f - fc) / 2.0f : 0.5f * fc;
        float ft = 2.0f * (0.25f - fc);
        unsigned h_len = filter_len_estimate(ft, as_s);
        unsigned m = (unsigned)std::ceil((float)(h_len - 1) / 4.0f);
        if (m < 3) m = 3;
        p.m.push_back((int)m);
    }
    return p;
}

// Equiripple (Parks-McClellan) half-band prototypes for stop-band 65 dB: the odd-index taps hf[1], hf[3], ...,
// hf[2m-1] of the 4m+1 tap prototype (even taps are zero, hf[2m] ~ 0.5, symmetric).  These three designs are
// the only ones msresamp ever instantiates at As = 60 (m = 3, 5, 10); values are the converged Remez solution
// (checked bit-for-bit against the reference binary in tests/test_design.py).
// (IEEE-754 single bit patterns; decimal values in the comments)
static const uint32_t kHb3[3] = {0x3c18fdc4u, 0xbd6aff07u, 0x3e98aab6u};   // 0.00933784619, -0.0573721193, 0.298177421
static const uint32_t kHb5[5] = {0x3b25721du, 0xbc33cf23u, 0x3d04243cu, 0xbdaba2c5u, 0x3e9ed3eau};
//   0.00252450188, -0.0109746782, 0.0322611183, -0.083806552, 0.310210526
static const uint32_t kHb10[10] = {0xba3b0279u, 0x3ae57b8du, 0xbb74c970u, 0x3be61d5au, 0xbc4791eeu,
                                   0x3ca4ad91u, 0xbd052c32u, 0x3d5c3bc8u, 0xbdcc95e4u, 0x3ea1e41fu};
//   -0.000713385234, 0.00175081345, -0.00373515114, 0.00702254195, -0.0121807884,
//    0.0201022942, -0.0325128511, 0.0537679493, -0.0998952687, 0.316193551

bool halfband_branch_taps(int m, float as_db, float *h1) {
    if (std::fabs(as_db - 65.0f) > 1e-3f) return false;
    const uint32_t *odd = (m == 3) ? kHb3 : (m == 5) ? kHb5 : (m == 10) ? kHb10 : nullptr;
    if (!odd) return false;
    // branch taps = 2 * prototype odd taps, laid out oldest-sample-first; the prototype is symmetric
    for (int j = 0; j < m; ++j) {
        float v;
        std::memcpy(&v, &odd[j], sizeof(v));
        h1[j] = 2.0f * v;
        h1[2 * m - 1 - j] = 2.0f * v;
    }
    return true;
}

// ---- multi-stage resampler plan ----------------------------------------------------------------------------------
ResamplerPlan make_resampler_plan(float rate, float as_db) {
    ResamplerPlan p;
    p.rate = rate;
    p.interp = rate > 1.0f;
    float ra = rate;
    int S = 0;
    if (p.interp) { while (ra > 2.0f) { ++S; ra *= 0.5f; } }
    else          { while (ra < 0.5f) { ++S; ra *= 2.0f; } }
    p.num_halfband = S;
    p.rate_arb = ra;
    p.zeta = 1.0f / (float)(1u << S);
    p.hb = halfband_plan(S, as_db);
    p.arb_step = (uint32_t)roundf(16777216.0f / ra);
    float fc = 0.515f * ra;
    if (fc > 0.49f) fc = 0.49f;
    p.arb_fc = fc;
    const unsigned npfb = p.arb_npfb, m = p.arb_m, n = 2 * m * npfb + 1;
    std::vector<float> hf(n);
    kaiser_lowpass(n, fc / (float)npfb, as_db, 0.0f, hf.data());
    float gain = 0.0f;
    for (unsigned i = 0; i < n; ++i) gain += hf[i];
    gain = (float)npfb / gain;
    p.pfb.resize((size_t)npfb * 2 * m);
    for (unsigned a = 0; a < npfb; ++a)
        for (unsigned k = 0; k < 2 * m; ++k) p.pfb[(size_t)a * 2 * m + k] = hf[a + k * npfb] * gain;
    return p;
}

float ResamplerPlan::delay() const {
    // group delay in input samples: half-band stages then the arbitrary stage (informational only)
    float d = 0.f;
    for (int g = 0; g < num_halfband; ++g) d = 2.0f * d + (float)(2 * hb.m[g]) - 0.5f;
    return d + (float)arb_m;
}

uint64_t arb_output_count(uint32_t step, uint64_t phase, uint64_t n_in, uint64_t *phase_after) {
    const uint64_t span = n_in << 24;
    uint64_t cnt = 0;
    if (phase < span) cnt = (span - phase + step - 1) / step;
    if (phase_after) *phase_after = phase + cnt * step - span;
    return cnt;
}

// ---- demodulator filters -------------------------------------------------------------------------------------------
void am_dc_block_taps(unsigned m, float as_db, float *h) {
    // Kaiser-window notch at DC: h = delta - w / sum(w)
    const unsigned n = 2 * m + 1;
    float beta = kaiser_beta_for_atten(as_db), scale = 0.f;
    for (unsigned i = 0; i < n; ++i) {
        float w = kaiser_window(i, n, beta);
        h[i] = -w;
        scale += w;
    }
    for (unsigned i = 0; i < n; ++i) h[i] /= scale;
    h[m] += 1.0f;
}

void hilbert_quadrature_taps(unsigned m, float as_db, float *hq) {
    // half-band Kaiser low-pass shifted by fs/4; the quadrature branch keeps its odd taps (imaginary parts)
    const unsigned n = 4 * m + 1;
    std::vector<float> h(n);
    kaiser_lowpass(n, 0.25f, as_db, 0.0f, h.data());
    for (unsigned i = 0; i < n; ++i) {
        float t = (float)i - (float)(n - 1) / 2.0f;
        h[i] = h[i] * sinf((float)(0.5 * kPi * (double)t));
    }
    unsigned j = 0;
    for (unsigned i = 1; i < n; i += 2) hq[j++] = h[n - i - 1];
}

void butterworth_lowpass_sos(unsigned order, float fc, std::vector<Sos> &sos) {
    // analog Butterworth poles on the unit circle, bilinear transform with pre-warp tan(pi fc), conjugate pairs
    // expanded into second-order sections, overall gain split evenly over the sections' numerators
    typedef std::complex<double> cd;
    const unsigned r = order % 2, L = (order - r) / 2;
    std::vector<cd> pa;
    for (unsigned i = 0; i < L; ++i) {
        double theta = (double)(2 * (i + 1) + order - 1) * kPi / (double)(2 * order);
        pa.push_back(std::polar(1.0, theta));
        pa.push_back(std::polar(1.0, -theta));
    }
    if (r) pa.push_back(cd(-1.0, 0.0));
    const double mw = std::tan(kPi * (double)fc);
    std::vector<cd> pd(order);
    cd G(1.0, 0.0);
    for (unsigned i = 0; i < order; ++i) {
        cd pm = pa[i] * mw;
        pd[i] = (1.0 + pm) / (1.0 - pm);
        G *= (1.0 - pd[i]) / (1.0 - cd(-1.0, 0.0));
    }
    const double k = std::pow(G.real(), 1.0 / (double)(L + r));
    sos.clear();
    for (unsigned i = 0; i < L; ++i) {
        cd p0 = -pd[2 * i], p1 = -pd[2 * i + 1];
        Sos s;
        s.a[0] = 1.0f; s.a[1] = (float)(p0 + p1).real(); s.a[2] = (float)(p0 * p1).real();
        s.b[0] = (float)k; s.b[1] = (float)(2.0 * k); s.b[2] = (float)k;      // zeros at z = -1 (twice)
        sos.push_back(s);
    }
    if (r) {
        cd p0 = -pd[order - 1];
        Sos s;
        s.a[0] = 1.0f; s.a[1] = (float)p0.real(); s.a[2] = 0.f;
        s.b[0] = (float)k; s.b[1] = (float)k; s.b[2] = 0.f;
        sos.push_back(s);
    }
}


// ---- Chebyshev-II band-pass, order 5 (liquid_iirdes as iirfilt_crcf_create_prototype calls it, ModemFMStereo.cpp:136) --------
namespace {
struct cf { float re, im; };
inline float bits_f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
// cheby2_azpkf(5, 10^(-60/20)): analog zeros and poles of the prototype (its gain is not used).  Independent of the rates, so carried as the bit
// patterns the reference computes (as the half-band prototypes are).
const uint32_t kCheby2Za[4][2] = {{0x00000000u, 0x3f869650u}, {0x80000000u, 0xbf869650u}, {0x00000000u, 0x3fd9c440u}, {0x80000000u, 0xbfd9c440u}};
const uint32_t kCheby2Pa[5][2] = {{0xbdf46b57u, 0x3ecebecbu}, {0xbdf46b57u, 0xbecebecbu}, {0xbeb15f7fu, 0x3e8dac18u}, {0xbeb15f7fu, 0xbe8dac18u},
                                  {0xbeeb0b3bu, 0x00000000u}};

// liquid_cplxpair(z, n, 1e-6, p): conjugate pairs first (in order of discovery), then the real values; then
// liquid_cplxpair_cleanup: exact conjugates with the negative imaginary part first, pairs sorted by real part, reals sorted
void cplxpair(const cf *z, unsigned n, float tol, cf *p) {
    std::vector<unsigned char> paired(n, 0);
    unsigned k = 0, num_pairs = 0;
    for (unsigned i = 0; i < n; ++i) {
        if (paired[i] || fabsf(z[i].im) < tol) continue;
        for (unsigned j = 0; j < n; ++j) {
            if (j == i || paired[j] || fabsf(z[j].im) < tol) continue;
            if (fabsf(z[i].im + z[j].im) < tol && fabsf(z[i].re - z[j].re) < tol) {
                p[k++] = z[i]; p[k++] = z[j];
                paired[i] = paired[j] = 1;
                ++num_pairs;
                break;
            }
        }
    }
    for (unsigned i = 0; i < n; ++i) if (!paired[i]) { p[k++] = z[i]; paired[i] = 1; }
    for (unsigned i = 0; i < num_pairs; ++i) {
        if (!(p[2 * i].im < 0)) p[2 * i].im = -p[2 * i].im;
        p[2 * i + 1].re = p[2 * i].re; p[2 * i + 1].im = -p[2 * i].im;
    }
    for (unsigned i = 0; i < num_pairs; ++i)
        for (unsigned j = num_pairs - 1; j > i; --j)
            if (p[2 * j].re < p[2 * (j - 1)].re) { std::swap(p[2 * j], p[2 * (j - 1)]); std::swap(p[2 * j + 1], p[2 * (j - 1) + 1]); }
    for (unsigned i = 2 * num_pairs; i < n; ++i)
        for (unsigned j = n - 1; j > i; --j)
            if (p[j].re < p[j - 1].re) std::swap(p[j], p[j - 1]);
}
}  // namespace

void cheby2_bandpass5_sos(float fc, float f0, std::vector<Sos> &sos) {
    const unsigned n = 5, nza = 4;
    cf za[5], pa[5];
    for (unsigned i = 0; i < nza; ++i) { za[i].re = bits_f(kCheby2Za[i][0]); za[i].im = bits_f(kCheby2Za[i][1]); }
    for (unsigned i = 0; i < n; ++i) { pa[i].re = bits_f(kCheby2Pa[i][0]); pa[i].im = bits_f(kCheby2Pa[i][1]); }
    // iirdes_freqprewarp(BANDPASS): |(cos(2 pi fc) - cos(2 pi f0)) / sin(2 pi fc)|, the angles narrowed to float first
    const double two_pi = 6.283185307179586;
    float sn, cs;
    sincosf((float)((double)fc * two_pi), &sn, &cs);
    volatile float c0 = cosf((float)((double)f0 * two_pi));
    volatile float num = cs - c0;
    volatile float mq = num / sn;
    const float m = fabsf(mq);
    // bilinear_zpkf: zd = (1 + m za) / (1 - m za) and the same for the poles, evaluated in double from the float products and
    // narrowed; the gain G *= (1 - pd) / (1 - zd) in double from the narrowed values, G itself kept in float between factors
    cf zd[10], pd[10];
    float Gr = 1.0f, Gi = 0.0f;                                      // liquid_iirdes hands the transform its nominal gain k0 = 1, not ka
    for (unsigned i = 0; i < n; ++i) {
        if (i < nza) {
            volatile float zmr = za[i].re * m, zmi = za[i].im * m;
            const double zr = (double)zmr, zi = (double)zmi;
            const double dr = 1.0 - zr, nr = zr + 1.0, zi2 = zi * zi;
            const double den = dr * dr + zi2;
            zd[i].re = (float)((nr * dr - zi2) / den);
            zd[i].im = (float)((zi + zi) / den);
        } else { zd[i].re = -1.0f; zd[i].im = 0.0f; }
        {
            volatile float pmr = pa[i].re * m, pmi = pa[i].im * m;
            const double pr = (double)pmr, pi = (double)pmi;
            const double dr = 1.0 - pr, nr = pr + 1.0, pi2 = pi * pi;
            const double den = dr * dr + pi2;
            pd[i].re = (float)((nr * dr - pi2) / den);
            pd[i].im = (float)((pi + pi) / den);
        }
        const double zr = (double)zd[i].re, zi = (double)zd[i].im, pr = (double)pd[i].re, pi = (double)pd[i].im;
        const double omz = 1.0 - zr, den = omz * omz + zi * zi;
        const double qr = ((1.0 - pr) * omz + pi * zi) / den;
        const double qi = ((zr - 1.0) * pi - (pr - 1.0) * zi) / den;
        const double gr = (double)Gr, gi = (double)Gi;
        Gr = (float)(gr * qr - gi * qi);
        Gi = (float)(gi * qr + gr * qi);
    }
    // iirdes_dzpk_lp2bp: every zero / pole z becomes 0.5 (c0 (1 + z) +- csqrt(c0^2 (1 + z)^2 - 4 z)), all in float
    const float cc = cosf((float)((double)f0 * two_pi));
    volatile float cc2 = cc * cc;
    cf zt[10], pt[10];
    for (int which = 0; which < 2; ++which) {
        const cf *src = which ? pd : zd;
        cf *dst = which ? pt : zt;
        for (unsigned i = 0; i < n; ++i) {
            const float zr = src[i].re, zi = src[i].im;
            volatile float t0r = zr + 1.0f;
            volatile float a = cc2 * t0r, b = cc2 * zi;              // c0^2 t0
            volatile float ar = a * t0r, bz = b * zi, az = a * zi, tb = t0r * b;
            volatile float re2 = ar - bz, im2 = az + tb;             // c0^2 t0^2
            volatile float fzr = 4.0f * zr, fzi = 4.0f * zi;
            volatile float argr = re2 - fzr, argi = im2 - fzi;
            const std::complex<float> sq = std::sqrt(std::complex<float>(argr, argi));
            volatile float hr = cc * t0r, hi = zi * cc;
            volatile float s1r = sq.real() + hr, s1i = sq.imag() + hi, s2r = hr - sq.real(), s2i = hi - sq.imag();
            dst[2 * i].re = s1r * 0.5f; dst[2 * i].im = s1i * 0.5f;
            dst[2 * i + 1].re = s2r * 0.5f; dst[2 * i + 1].im = s2i * 0.5f;
        }
    }
    // iirdes_dzpk2sosf: pair, then per pair  a = {1, -(re0 + re1), re0 re1 - im0 im1}, the same for the zeros, and the gain
    // kd^(1 / sections) on every numerator
    cf zp[10], pp[10];
    cplxpair(zt, 2 * n, 1e-6f, zp);
    cplxpair(pt, 2 * n, 1e-6f, pp);
    const unsigned L = n;
    volatile float inv_l = 1.0f / (float)L;
    const float k = (float)pow((double)Gr, (double)inv_l);
    sos.clear();
    for (unsigned i = 0; i < L; ++i) {
        Sos s;
        volatile float sp = pp[2 * i].re + pp[2 * i + 1].re, pr = pp[2 * i].re * pp[2 * i + 1].re, pi = pp[2 * i + 1].im * pp[2 * i].im;
        volatile float sz = zp[2 * i].re + zp[2 * i + 1].re, zr = zp[2 * i].re * zp[2 * i + 1].re, zi = zp[2 * i + 1].im * zp[2 * i].im;
        volatile float a2 = pr - pi, b2 = zr - zi;
        s.a[0] = 1.0f; s.a[1] = -sp; s.a[2] = a2;
        volatile float b0 = 1.0f * k, b1 = -sz * k, b2k = b2 * k;
        s.b[0] = b0; s.b[1] = b1; s.b[2] = b2k;
        sos.push_back(s);
    }
}

// iirfiltsos_crcf_create (what iirfilt_crcf_create_sos builds each section with) normalises {b0, b1} by a reciprocal of a0 that the
// reference's binary takes from the SSE approximation instruction plus one Newton step, x = 2 x0 - a0 x0 x0 with x0 = rcpps(a0);
// b2, a1 and a2 are divided exactly.  For a0 = 1 that reciprocal is one ulp short of 1 on the CPUs at hand, which moves b0 and b1
// down by an ulp -- visible through a filter whose poles sit 0.002 inside the unit circle.  The same instruction is used here.
void sos_section_object(Sos &s) {
    const float a0 = s.a[0];
    float x0 = _mm_cvtss_f32(_mm_rcp_ps(_mm_set1_ps(a0)));
    volatile float ax = a0 * x0;
    volatile float axx = ax * x0;
    volatile float two = x0 + x0;
    volatile float r = two - axx;
    volatile float b0 = r * s.b[0], b1 = r * s.b[1];
    volatile float inv = 1.0f / a0;
    volatile float b2 = s.b[2] * inv, a1 = s.a[1] * inv, a2 = inv * s.a[2];
    s.b[0] = b0; s.b[1] = b1; s.b[2] = b2; s.a[0] = 1.0f; s.a[1] = a1; s.a[2] = a2;
}

void fm_deemphasis(int tau_us, int audio_rate, float *b2, float *a1) {
    const double f = (1.0 / (2.0 * kPi * double(tau_us) * 1e-6));
    double t = 1.0 / (2.0 * kPi * f);
    t = 1.0 / (2.0 * double(audio_rate) * tan(1.0 / (2.0 * double(audio_rate) * t)));
    const double tb = (1.0 + 2.0 * t * double(audio_rate));
    b2[0] = (float)(1.0 / tb); b2[1] = (float)(1.0 / tb);
    *a1 = (float)((1.0 - 2.0 * t * double(audio_rate)) / tb);
}

void pfb_analysis_tables(unsigned M, unsigned m, float as_db, std::vector<float> &hq, std::vector<float> &tw_re_im) {
    const unsigned p = 2 * m, h_len = 2 * M * m + 1;
    std::vector<float> h(h_len);
    kaiser_lowpass(h_len, 0.5f / (float)M, as_db, 0.0f, h.data());
    hq.assign((size_t)M * p, 0.f);
    for (unsigned c = 0; c < M; ++c)
        for (unsigned q = 0; q < p; ++q) hq[(size_t)c * p + q] = h[(M - 1 - c) + q * M];
    tw_re_im.assign((size_t)2 * M, 0.f);
    for (unsigned n = 0; n < M; ++n) {
        const double ph = -2.0 * 3.14159265358979323846 * (double)n / (double)M;
        tw_re_im[2 * n] = (float)cos(ph);
        tw_re_im[2 * n + 1] = (float)sin(ph);
    }
}

void pfb2_analysis_tables(unsigned M, unsigned m, float as_db, std::vector<float> &hq, std::vector<float> &tw_re_im) {
    const unsigned D = M / 2, taps = 4 * m, h_len = 2 * M * m + 1;
    std::vector<float> h(h_len);
    kaiser_lowpass(h_len, 1.0f / (float)M, as_db, 0.0f, h.data());
    // liquid normalises the prototype to M / sum(h) and scales the transform by 1 / M: folded into one factor here
    double sum = 0.0;
    for (float v : h) sum += (double)v;
    hq.assign((size_t)D * taps, 0.f);
    for (unsigned cm = 0; cm < D; ++cm)
        for (unsigned d = 0; d < taps; ++d) hq[(size_t)cm * taps + d] = (float)((double)h[cm + D * d] / sum);
    tw_re_im.assign((size_t)2 * M, 0.f);
    for (unsigned n = 0; n < M; ++n) {
        const double ph = 2.0 * 3.14159265358979323846 * (double)n / (double)M;
        tw_re_im[2 * n] = (float)cos(ph);
        tw_re_im[2 * n + 1] = (float)sin(ph);
    }
}

int optimal_channel_count(long long sample_rate) {
    if (sample_rate <= 500000) return 1;
    int count = (int)ceil((double)sample_rate / 500000.0);
    if (count % 2 == 1) count--;
    if (count < 2) count = 2;
    return count;
}

}  // namespace sdrb200
