/* oracle/liquid_port.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement ("port") of the hot path in plain C.
 *
 * The reference's arithmetic for this path lives in liquid-dsp 1.5.0 (pinned by the vendored header
 * /root/reference/external/liquid-dsp/include/liquid/liquid.h:59 LIQUID_VERSION "1.5.0"; the source is NOT in the tree).
 * This file restates the published algorithms, sample by sample and object by object the way liquid executes them
 * (SURVEY.md Appendix A documents the structure; line numbers below are the API text in liquid.h), plus the CubicSDR glue
 * of oracle/ref_chain.cpp.  It is validated against the reference's own binary (oracle/_ref, tests/test_port.py) and is the
 * fallback checker if the transplanted binary cannot be loaded on some machine.  Parity status: pinned (to the binary).
 *
 * Only tests/ and bench.py's CPU-baseline leg may use this file; the product never does.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PI 3.14159265358979323846
typedef struct { float re, im; } cf;

/* ---- filter design ---------------------------------------------------------------------------------------------- */
static float kaiser_beta(float As) {                       /* liquid.h:2122 kaiser_beta_As */
    As = fabsf(As);
    if (As > 50.0f) return 0.1102f * (As - 8.7f);
    if (As > 21.0f) return 0.5842f * powf(As - 21.0f, 0.4f) + 0.07886f * (As - 21.0f);
    return 0.0f;
}
static double bessel_i0(double z) {
    double q = 0.25 * z * z, t = 1.0, s = 1.0;
    for (int k = 1; k < 64; ++k) { t *= q / ((double)k * k); s += t; }
    return s;
}
static float kaiser_w(unsigned i, unsigned n, float beta) { /* liquid.h:6897 liquid_kaiser */
    float t = (float)i - (float)(n - 1) / 2, r = 2.0f * t / (float)(n - 1);
    double a = 1.0 - (double)r * r;
    if (a < 0) a = 0;
    return (float)(bessel_i0(beta * sqrt(a)) / bessel_i0(beta));
}
static float sincf_(float x) {
    if (fabsf(x) < 0.01f) return (float)(cos(PI * x / 2) * cos(PI * x / 4) * cos(PI * x / 8));
    return (float)(sin(PI * x) / (PI * x));
}
void port_firdes_kaiser(unsigned n, float fc, float As, float mu, float *h) {   /* liquid.h:2262 */
    float beta = kaiser_beta(As);
    for (unsigned i = 0; i < n; ++i) {
        float t = (float)i - (float)(n - 1) / 2 + mu;
        h[i] = sincf_(2.0f * fc * t) * kaiser_w(i, n, beta);
    }
}
/* Parks-McClellan half-band prototypes hf[1], hf[3], ... (liquid.h:2240 liquid_firdespm_halfband_as at As = 65):
 * the three designs msresamp uses at As = 60; decimal values as read from the reference binary (tests/test_port.py) */
static const float HB3[3] = {0.00933784619f, -0.0573721193f, 0.298177421f};
static const float HB5[5] = {0.00252450188f, -0.0109746782f, 0.0322611183f, -0.083806552f, 0.310210526f};
static const float HB10[10] = {-0.000713385234f, 0.00175081345f, -0.00373515114f, 0.00702254195f, -0.0121807884f,
                               0.0201022942f, -0.0325128511f, 0.0537679493f, -0.0998952687f, 0.316193551f};

/* ---- nco_crcf (liquid.h:8731-8860) ------------------------------------------------------------------ */
typedef struct { float tab[1024]; uint32_t theta, dtheta; } nco_t;
static void nco_init(nco_t *q) {
    for (int i = 0; i < 1024; ++i) q->tab[i] = sinf((float)i * 0.0061359233f);
    q->theta = q->dtheta = 0;
}
static void nco_set_frequency(nco_t *q, float f) {
    float x = (float)((double)f * 0.159154943091895);
    x -= truncf(x);
    if (x < 0) x += 1.0f;
    q->dtheta = (uint32_t)(int64_t)(x * 4294967296.0f);
}
static void nco_sincos(const nco_t *q, float *s, float *c) {
    uint32_t idx = (q->theta + 0x200000u) >> 22;
    *s = q->tab[idx]; *c = q->tab[(idx + 256) & 1023];
}
static cf nco_mix(const nco_t *q, cf x, int up) {
    float s, c; nco_sincos(q, &s, &c);
    if (!up) s = -s;
    cf y = {x.re * c - x.im * s, x.re * s + x.im * c};
    return y;
}

/* ---- resamp2 (liquid.h:4214, 4294), real coefficients, complex or real data ---------------------------------------- */
typedef struct { int m; float h1[20]; cf w0[20], w1[20]; } hb_t;   /* windows oldest-first, length 2m */
static void hb_init(hb_t *q, int m) {
    const float *odd = m == 3 ? HB3 : (m == 5 ? HB5 : HB10);
    memset(q, 0, sizeof *q);
    q->m = m;
    for (int j = 0; j < m; ++j) { q->h1[j] = 2.0f * odd[j]; q->h1[2 * m - 1 - j] = 2.0f * odd[j]; }
}
static void win_push(cf *w, int n, cf x) { memmove(w, w + 1, (n - 1) * sizeof(cf)); w[n - 1] = x; }
static cf hb_decim(hb_t *q, cf x0, cf x1) {
    const int n = 2 * q->m;
    win_push(q->w1, n, x0);
    cf y1 = {0, 0};
    for (int j = 0; j < n; ++j) { y1.re += q->h1[j] * q->w1[j].re; y1.im += q->h1[j] * q->w1[j].im; }
    win_push(q->w0, n, x1);
    cf y0 = q->w0[q->m - 1];
    cf y = {y0.re + y1.re, y0.im + y1.im};
    return y;
}
static void hb_interp(hb_t *q, cf x, cf *y) {
    const int n = 2 * q->m;
    win_push(q->w0, n, x); y[0] = q->w0[q->m - 1];
    win_push(q->w1, n, x);
    cf a = {0, 0};
    for (int j = 0; j < n; ++j) { a.re += q->h1[j] * q->w1[j].re; a.im += q->h1[j] * q->w1[j].im; }
    y[1] = a;
}

/* ---- resamp (fixed-point polyphase, liquid.h:4513, 4600) ----------------------------------------------- */
typedef struct { float *pfb; uint32_t step, phase; cf w[14]; } arb_t;
static void arb_init(arb_t *q, float rate, float fc, float As) {
    const unsigned npfb = 256, m = 7, n = 2 * m * npfb + 1;
    float *hf = (float *)malloc(n * sizeof(float));
    port_firdes_kaiser(n, fc / (float)npfb, As, 0.0f, hf);
    float g = 0;
    for (unsigned i = 0; i < n; ++i) g += hf[i];
    g = (float)npfb / g;
    q->pfb = (float *)malloc(npfb * 14 * sizeof(float));
    for (unsigned a = 0; a < npfb; ++a) for (unsigned k = 0; k < 14; ++k) q->pfb[a * 14 + k] = hf[a + k * npfb] * g;
    free(hf);
    q->step = (uint32_t)roundf(16777216.0f / rate);
    q->phase = 0;
    memset(q->w, 0, sizeof q->w);
}
static unsigned arb_execute(arb_t *q, cf x, cf *y) {
    win_push(q->w, 14, x);
    unsigned n = 0;
    while (q->phase < (1u << 24)) {
        const float *h = q->pfb + (q->phase >> 16) * 14;
        cf a = {0, 0};
        for (int k = 0; k < 14; ++k) { a.re += h[k] * q->w[13 - k].re; a.im += h[k] * q->w[13 - k].im; }
        y[n++] = a;
        q->phase += q->step;
    }
    q->phase -= (1u << 24);
    return n;
}

/* ---- msresamp (liquid.h:4736-4780) ------------------------------------------------------------------- */
typedef struct { int interp, S; hb_t hb[16]; arb_t arb; float zeta; cf buf[1 << 16]; unsigned buf_n; } msr_t;
static unsigned est_len(float df, float As) { return (unsigned)((As - 7.95f) / (14.26f * df)); }   /* liquid.h:2103 */
static void msr_init(msr_t *q, float r, float As) {
    memset(q, 0, sizeof *q);
    q->interp = r > 1.0f;
    float ra = r;
    if (q->interp) while (ra > 2.0f) { q->S++; ra *= 0.5f; }
    else while (ra < 0.5f) { q->S++; ra *= 2.0f; }
    float fc = 0.4f;
    for (int i = 0; i < q->S; ++i) {
        fc = (i == 1) ? (0.5f - fc) / 2.0f : 0.5f * fc;
        float ft = 2.0f * (0.25f - fc);
        unsigned hl = est_len(ft, As + 5.0f);
        int m = (int)ceilf((float)(hl - 1) / 4.0f);
        if (m < 3) m = 3;
        hb_init(&q->hb[i], m);
    }
    q->zeta = 1.0f / (float)(1u << q->S);
    float afc = 0.515f * ra;
    if (afc > 0.49f) afc = 0.49f;
    arb_init(&q->arb, ra, afc, As);
}
static void msr_free(msr_t *q) { free(q->arb.pfb); }
static cf ms2_decim(msr_t *q, cf *b) {                       /* msresamp2 decim: stage S-1 first, x zeta at the end */
    unsigned n = 1u << q->S;
    for (int s = 0; s < q->S; ++s) {
        int g = q->S - 1 - s;
        n >>= 1;
        for (unsigned i = 0; i < n; ++i) b[i] = hb_decim(&q->hb[g], b[2 * i], b[2 * i + 1]);
    }
    cf y = {b[0].re * q->zeta, b[0].im * q->zeta};
    return y;
}
static unsigned msr_execute(msr_t *q, const cf *x, unsigned nx, cf *y) {
    unsigned ny = 0;
    if (!q->interp) {
        cf tmp[1 << 16];
        for (unsigned i = 0; i < nx; ++i) {
            q->buf[q->buf_n++] = x[i];
            if (q->buf_n == (1u << q->S)) {
                memcpy(tmp, q->buf, q->buf_n * sizeof(cf));
                cf z = ms2_decim(q, tmp);
                ny += arb_execute(&q->arb, z, y + ny);
                q->buf_n = 0;
            }
        }
    } else {
        for (unsigned i = 0; i < nx; ++i) {
            cf a[4];
            unsigned na = arb_execute(&q->arb, x[i], a);
            for (unsigned j = 0; j < na; ++j) {                /* msresamp2 interp: stage 0 first */
                cf b0[256], b1[256];
                b0[0] = a[j];
                unsigned n = 1;
                for (int s = 0; s < q->S; ++s) {
                    for (unsigned k = 0; k < n; ++k) hb_interp(&q->hb[s], b0[k], &b1[2 * k]);
                    n <<= 1;
                    memcpy(b0, b1, n * sizeof(cf));
                }
                memcpy(y + ny, b0, n * sizeof(cf));
                ny += n;
            }
        }
    }
    return ny;
}

/* ---- demodulator filters -------------------------------------------------------------------------------------------------- */
static void am_notch_taps(float *h) {                        /* firfilt_rrrf_create_dc_blocker(25,30): liquid.h:2825 */
    float beta = kaiser_beta(30.0f), sc = 0;
    for (int i = 0; i < 51; ++i) { float w = kaiser_w(i, 51, beta); h[i] = -w; sc += w; }
    for (int i = 0; i < 51; ++i) h[i] /= sc;
    h[25] += 1.0f;
}
static void hilbert_q(float *hq) {                           /* firhilbf_create(5,90): liquid.h:3062 */
    float h[21];
    port_firdes_kaiser(21, 0.25f, 90.0f, 0.0f, h);
    for (int i = 0; i < 21; ++i) h[i] *= sinf((float)(0.5 * PI * (double)(i - 10)));
    int j = 0;
    for (int i = 1; i < 21; i += 2) hq[j++] = h[21 - i - 1];
}
static void butter6(float sos[3][5]) {                       /* iirfilt_crcf_create_lowpass(6,0.25): liquid.h:3402 */
    double k = 1.0, pr[3], pa2[3];
    double G = 1.0;
    for (int i = 0; i < 3; ++i) {
        double th = (double)(2 * (i + 1) + 5) * PI / 12.0, m = tan(PI * 0.25);
        /* p = (1 + m e^{j th}) / (1 - m e^{j th}) and its conjugate */
        double ar = 1 + m * cos(th), ai = m * sin(th), br = 1 - m * cos(th), bi = -m * sin(th);
        double d = br * br + bi * bi, pre = (ar * br + ai * bi) / d, pim = (ai * br - ar * bi) / d;
        pr[i] = -2 * pre; pa2[i] = pre * pre + pim * pim;
        double gr = (1 - pre), gi = -pim;                      /* (1-p)(1-p*) / ((1+1)(1+1)) */
        G *= (gr * gr + gi * gi) / 4.0;
    }
    k = pow(G, 1.0 / 3.0);
    for (int i = 0; i < 3; ++i) { sos[i][0] = (float)k; sos[i][1] = (float)(2 * k); sos[i][2] = (float)k; sos[i][3] = (float)pr[i]; sos[i][4] = (float)pa2[i]; }
}

/* ---- one channel: DemodulatorPreThread + Modem + metering (same glue as oracle/ref_chain.cpp) ----------------------------- */
enum { M_FM, M_NBFM, M_AM, M_USB, M_LSB };
typedef struct {
    int modem; long long fs; int bw, fa;
    nco_t shifter, ssb; long long shift; msr_t *iq, *au;
    cf fm_prev; float am_h[51], am_w[51], hq[10]; float sos[3][5]; cf sv[3][2];
    cf hw[21]; float ceil_, ma, maa;
} pchan;
static int modem_id(const char *n) {
    return !strcmp(n, "FM") ? M_FM : !strcmp(n, "NBFM") ? M_NBFM : !strcmp(n, "AM") ? M_AM : !strcmp(n, "USB") ? M_USB : !strcmp(n, "LSB") ? M_LSB : -1;
}
void *port_chan_create(const char *modem, long long fs, long long bw, int fa) {
    int m = modem_id(modem);
    if (m < 0) return 0;
    pchan *c = (pchan *)calloc(1, sizeof *c);
    c->modem = m; c->fs = fs; c->fa = fa;
    if (bw < 500) bw = 500;
    if ((m == M_USB || m == M_LSB) && (bw % 2)) bw += 1;
    c->bw = (int)bw;
    nco_init(&c->shifter); nco_init(&c->ssb);
    nco_set_frequency(&c->ssb, (float)((2.0 * PI) * 0.25));
    c->iq = (msr_t *)malloc(sizeof(msr_t)); c->au = (msr_t *)malloc(sizeof(msr_t));
    msr_init(c->iq, (float)((double)c->bw / (double)fs), 60.0f);
    msr_init(c->au, (float)((double)fa / (double)c->bw), 60.0f);
    am_notch_taps(c->am_h); hilbert_q(c->hq); butter6(c->sos);
    c->ceil_ = c->ma = c->maa = 1.0f;
    return c;
}
void port_chan_destroy(void *h) { pchan *c = (pchan *)h; msr_free(c->iq); msr_free(c->au); free(c->iq); free(c->au); free(c); }

/* returns 1 if the block was skipped (|shift| > 0.75 fs) */
int port_chan_process(void *h, const cf *x, unsigned n, long long shift, cf *iq_out, unsigned *n_iq, float *audio, unsigned *n_audio,
                      double *level_db, float *peak) {
    pchan *c = (pchan *)h;
    *n_iq = *n_audio = 0;
    long long lim = (long long)(int)((double)(c->fs / 2) * 1.5);
    if (shift != c->shift) {
        c->shift = shift;
        if (llabs(shift) <= lim) nco_set_frequency(&c->shifter, (float)((2.0 * PI) * ((double)llabs(shift) / (double)c->fs)));
    }
    if (llabs(c->shift) > lim) return 1;
    cf *mix = (cf *)malloc(n * sizeof(cf));
    for (unsigned i = 0; i < n; ++i) {
        if (c->shift) { mix[i] = nco_mix(&c->shifter, x[i], c->shift < 0); c->shifter.theta += c->shifter.dtheta; }
        else mix[i] = x[i];
    }
    unsigned ni = msr_execute(c->iq, mix, n, iq_out);
    free(mix);
    *n_iq = ni;
    if (!ni) return 0;
    float *d = (float *)malloc(ni * sizeof(float));
    int agc = 0;
    for (unsigned i = 0; i < ni; ++i) {
        cf s = iq_out[i];
        if (c->modem == M_FM || c->modem == M_NBFM) {
            float re = c->fm_prev.re * s.re + c->fm_prev.im * s.im, im = c->fm_prev.re * s.im - c->fm_prev.im * s.re;
            d[i] = atan2f(im, re) * 0.31830987f;
            c->fm_prev = s;
        } else if (c->modem == M_AM) {
            memmove(c->am_w, c->am_w + 1, 50 * sizeof(float));
            c->am_w[50] = sqrtf(s.re * s.re + s.im * s.im);
            float a = 0;
            for (int k = 0; k < 51; ++k) a += c->am_h[k] * c->am_w[50 - k];
            d[i] = a; agc = 1;
        } else {
            c->ssb.theta += c->ssb.dtheta;
            cf v = nco_mix(&c->ssb, s, c->modem == M_LSB);
            for (int k = 0; k < 3; ++k) {
                const float *q = c->sos[k];
                cf v0 = {v.re - q[3] * c->sv[k][0].re - q[4] * c->sv[k][1].re, v.im - q[3] * c->sv[k][0].im - q[4] * c->sv[k][1].im};
                cf y = {q[0] * v0.re + q[1] * c->sv[k][0].re + q[2] * c->sv[k][1].re, q[0] * v0.im + q[1] * c->sv[k][0].im + q[2] * c->sv[k][1].im};
                c->sv[k][1] = c->sv[k][0]; c->sv[k][0] = v0; v = y;
            }
            cf w = nco_mix(&c->ssb, v, c->modem == M_USB);
            memmove(c->hw, c->hw + 1, 20 * sizeof(cf));
            c->hw[20] = w;
            float yq = 0;
            for (int j = 0; j < 10; ++j) yq += c->hq[j] * c->hw[20 - 19 + 2 * j].im;
            float yi = c->hw[10].re;
            d[i] = (c->modem == M_USB) ? yi - yq : yi + yq;
            agc = 1;
        }
    }
    if (agc) {                                               /* ModemAnalog::buildAudioOutput ModemAnalog.cpp:70-86 */
        c->ma = c->ma + (c->ceil_ - c->ma) * 0.025f;
        c->maa = c->maa + (c->ma - c->maa) * 0.025f;
        c->ceil_ = 0;
        for (unsigned i = 0; i < ni; ++i) if (d[i] > c->ceil_) c->ceil_ = d[i];
        float g = 0.5f / c->maa;
        for (unsigned i = 0; i < ni; ++i) d[i] *= g;
    }
    cf *dc_ = (cf *)malloc(ni * sizeof(cf)), *ac = (cf *)malloc(((size_t)(ni * ((double)c->fa / c->bw)) + 1024) * sizeof(cf));
    for (unsigned i = 0; i < ni; ++i) { dc_[i].re = d[i]; dc_[i].im = 0; }
    unsigned na = msr_execute(c->au, dc_, ni, ac);
    double acc = 0; float pk = 0;
    for (unsigned i = 0; i < na; ++i) { audio[i] = ac[i].re; float a = fabsf(audio[i]); if (a > pk) pk = a; if (agc) acc += a; }
    if (!agc) for (unsigned i = 0; i < ni; ++i) acc += sqrt((double)iq_out[i].re * iq_out[i].re + (double)iq_out[i].im * iq_out[i].im);
    *n_audio = na; *peak = pk;
    double lin = na ? acc / (double)(agc ? na : ni) : 0;
    if (lin <= 1e-20) lin = 1e-20;
    *level_db = na ? 20.0 * log10(lin) : 0;
    free(d); free(dc_); free(ac);
    return 0;
}

/* ---- DC blocker (SDRPostThread.cpp:29,284): float32 direct-form II, exactly the reference's operation order --------------- */
typedef struct { float vr, vi; } pdc;
void *port_dc_create(void) { return calloc(1, sizeof(pdc)); }
void port_dc_destroy(void *h) { free(h); }
void port_dc_block(void *h, const cf *x, unsigned n, cf *y) {
    pdc *q = (pdc *)h;
    const float c = 1.0f - 0.0005f;
    for (unsigned i = 0; i < n; ++i) {
        volatile float tr = c * q->vr, ti = c * q->vi;
        volatile float nr = x[i].re + tr, ni = x[i].im + ti;
        y[i].re = nr - q->vr; y[i].im = ni - q->vi;
        q->vr = nr; q->vi = ni;
    }
}

/* ---- spectrum (SpectrumVisualProcessor.cpp:388-576 non-view; radix-2 forward FFT as liquid's fft_execute for 2^k) ----------- */
typedef struct { unsigned fft, N; double *ma, *maa; double c_ma, c_maa, f_ma, f_maa; float rate; } pspec;
void *port_spectrum_create(unsigned fft) {
    pspec *s = (pspec *)calloc(1, sizeof *s);
    s->fft = fft; s->N = 2 * fft;
    s->ma = (double *)calloc(s->N, sizeof(double)); s->maa = (double *)calloc(s->N, sizeof(double));
    s->c_ma = s->c_maa = 100.0; s->rate = 0.65f;
    return s;
}
void port_spectrum_destroy(void *h) { pspec *s = (pspec *)h; free(s->ma); free(s->maa); free(s); }
static void fft_radix2(cf *x, unsigned n) {                  /* in-place decimation in time, e^{-j 2 pi k n / N}, unnormalised */
    for (unsigned i = 1, j = 0; i < n; ++i) {
        unsigned bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { cf t = x[i]; x[i] = x[j]; x[j] = t; }
    }
    for (unsigned len = 2; len <= n; len <<= 1) {
        for (unsigned i = 0; i < n; i += len)
            for (unsigned k = 0; k < len / 2; ++k) {
                double a = -2.0 * PI * k / len;
                float wr = (float)cos(a), wi = (float)sin(a);
                cf u = x[i + k], v = x[i + k + len / 2];
                cf t = {v.re * wr - v.im * wi, v.re * wi + v.im * wr};
                x[i + k].re = u.re + t.re; x[i + k].im = u.im + t.im;
                x[i + k + len / 2].re = u.re - t.re; x[i + k + len / 2].im = u.im - t.im;
            }
    }
}
void port_spectrum_frame(void *h, const cf *frame, float *points, unsigned char *line, double *mags) {
    pspec *s = (pspec *)h;
    const unsigned N = s->N;
    cf *X = (cf *)malloc(N * sizeof(cf));
    memcpy(X, frame, N * sizeof(cf));
    fft_radix2(X, N);
    float ceil_ = 0, floor_ = 1;
    for (unsigned i = 0; i < N; ++i) {
        unsigned k = (i + N / 2) % N;
        double m = (double)sqrtf(X[k].re * X[k].re + X[k].im * X[k].im);
        if (mags) mags[i] = m;
        s->maa[i] += (s->ma[i] - s->maa[i]) * s->rate;
        s->ma[i] += (m - s->ma[i]) * s->rate;
        if (s->maa[i] > ceil_) ceil_ = (float)s->maa[i];
        if (s->maa[i] < floor_) floor_ = (float)s->maa[i];
    }
    s->c_ma += (ceil_ - s->c_ma) * 0.05; s->c_maa += (s->c_ma - s->c_maa) * 0.05;
    s->f_ma += (floor_ - s->f_ma) * 0.05; s->f_maa += (s->f_ma - s->f_maa) * 0.05;
    for (unsigned x = 0; x < s->fft; ++x) {
        double acc = (2 * x > 0 ? s->maa[2 * x] : s->f_maa) + s->maa[2 * x + 1];
        float y = (float)(log10(acc / 2.0 + 0.25 - (s->f_maa - 0.75)) / log10((s->c_maa + 0.25) - (s->f_maa - 0.75)));
        points[x] = y;
        float wv = y < 0 ? 0 : (y > 0.99 ? (float)0.99 : y);
        line[x] = (unsigned char)floor(wv * 255.0);
    }
    free(X);
}
