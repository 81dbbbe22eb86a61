/* oracle/liquid_ms.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Minimal prototypes (written here, not copied) for the liquid-dsp 1.5.0 entry points CubicSDR's hot path
 * calls, as exported by the reference's vendored MinGW archive after the oracle/build_ref.py transplant.
 * The objects use the Microsoft x64 calling convention, hence LQ on every prototype; a complex sample is an
 * 8-byte aggregate passed/returned by value in an integer register.
 * API text: /root/reference/external/liquid-dsp/include/liquid/liquid.h (line numbers beside each group).
 */
#ifndef ORACLE_LIQUID_MS_H
#define ORACLE_LIQUID_MS_H
#ifdef __cplusplus
extern "C" {
#endif

#define LQ __attribute__((ms_abi))
typedef struct { float real, imag; } lq_cf;        /* liquid.h:153 liquid_float_complex */

LQ const char *liquid_libversion(void);            /* liquid.h:66 */

/* filter design helpers, liquid.h:2103-2122, 2240, 2262, 6822, 6897 */
LQ unsigned int estimate_req_filter_len(float df, float as);
LQ float estimate_req_filter_df(float as, unsigned int n);
LQ float kaiser_beta_As(float as);
LQ int liquid_firdespm_halfband_as(unsigned int m, float as, float *h);
LQ int liquid_firdes_kaiser(unsigned int n, float fc, float as, float mu, float *h);
LQ float liquid_besseli0f(float z);
LQ float liquid_kaiser(unsigned int i, unsigned int wlen, float beta);

/* nco_crcf, liquid.h:8731-8860 */
typedef struct nco_crcf_s *nco_crcf;
LQ nco_crcf nco_crcf_create(int type);
LQ int nco_crcf_destroy(nco_crcf q);
LQ int nco_crcf_reset(nco_crcf q);
LQ float nco_crcf_get_frequency(nco_crcf q);
LQ int nco_crcf_set_frequency(nco_crcf q, float dtheta);
LQ float nco_crcf_get_phase(nco_crcf q);
LQ int nco_crcf_set_phase(nco_crcf q, float phi);
LQ int nco_crcf_step(nco_crcf q);
LQ int nco_crcf_mix_up(nco_crcf q, lq_cf x, lq_cf *y);
LQ int nco_crcf_mix_down(nco_crcf q, lq_cf x, lq_cf *y);
LQ int nco_crcf_mix_block_up(nco_crcf q, lq_cf *x, lq_cf *y, unsigned int n);
LQ int nco_crcf_mix_block_down(nco_crcf q, lq_cf *x, lq_cf *y, unsigned int n);
/* phase-locked-loop entry points used by ModemFMStereo.cpp:143-145, 239-250 */
LQ int nco_crcf_cexpf(nco_crcf q, lq_cf *y);
LQ int nco_crcf_pll_set_bandwidth(nco_crcf q, float bw);
LQ int nco_crcf_pll_step(nco_crcf q, float dphi);

/* resamp2 / resamp / msresamp2 / msresamp, liquid.h:4214, 4294, 4513, 4600, 4663, 4736-4780 */
typedef struct resamp2_crcf_s *resamp2_crcf;
LQ resamp2_crcf resamp2_crcf_create(unsigned int m, float f0, float as);
LQ int resamp2_crcf_destroy(resamp2_crcf q);
LQ void resamp2_crcf_decim_execute(resamp2_crcf q, lq_cf *x, lq_cf *y);
typedef struct resamp_crcf_s *resamp_crcf;
LQ resamp_crcf resamp_crcf_create(float rate, unsigned int m, float fc, float as, unsigned int npfb);
LQ int resamp_crcf_destroy(resamp_crcf q);
LQ int resamp_crcf_execute(resamp_crcf q, lq_cf x, lq_cf *y, unsigned int *nw);
typedef struct msresamp_crcf_s *msresamp_crcf;
LQ msresamp_crcf msresamp_crcf_create(float r, float as);
LQ int msresamp_crcf_destroy(msresamp_crcf q);
LQ float msresamp_crcf_get_delay(msresamp_crcf q);
LQ int msresamp_crcf_execute(msresamp_crcf q, lq_cf *x, unsigned int nx, lq_cf *y, unsigned int *ny);
typedef struct msresamp_rrrf_s *msresamp_rrrf;
LQ msresamp_rrrf msresamp_rrrf_create(float r, float as);
LQ int msresamp_rrrf_destroy(msresamp_rrrf q);
LQ float msresamp_rrrf_get_delay(msresamp_rrrf q);
LQ int msresamp_rrrf_execute(msresamp_rrrf q, float *x, unsigned int nx, float *y, unsigned int *ny);

/* msresamp_cccf (ModemCW.cpp:125,158): complex samples, complex taps */
typedef struct msresamp_cccf_s *msresamp_cccf;
LQ msresamp_cccf msresamp_cccf_create(float r, float as);
LQ int msresamp_cccf_destroy(msresamp_cccf q);
LQ int msresamp_cccf_execute(msresamp_cccf q, lq_cf *x, unsigned int nx, lq_cf *y, unsigned int *ny);

/* freqdem, liquid.h:8247-8275 */
typedef struct freqdem_s *freqdem;
LQ freqdem freqdem_create(float kf);
LQ int freqdem_destroy(freqdem q);
LQ int freqdem_demodulate_block(freqdem q, lq_cf *r, unsigned int n, float *m);

/* firfilt_rrrf, liquid.h:2825, 2873, 2888 */
typedef struct firfilt_rrrf_s *firfilt_rrrf;
LQ firfilt_rrrf firfilt_rrrf_create_dc_blocker(unsigned int m, float as);
LQ firfilt_rrrf firfilt_rrrf_create(float *h, unsigned int n);          /* ModemFMStereo.cpp:122-123 */
LQ int firfilt_rrrf_destroy(firfilt_rrrf q);
LQ int firfilt_rrrf_push(firfilt_rrrf q, float x);
LQ int firfilt_rrrf_execute(firfilt_rrrf q, float *y);

/* iirfilt_crcf, liquid.h:3402, 3414, 3441, 3452 */
typedef struct iirfilt_crcf_s *iirfilt_crcf;
LQ iirfilt_crcf iirfilt_crcf_create_lowpass(unsigned int order, float fc);
LQ iirfilt_crcf iirfilt_crcf_create_dc_blocker(float alpha);
/* liquid.h:3387 (ModemFMStereo.cpp:136): ftype CHEBY2 = 2, btype BANDPASS = 2, format SOS = 0 */
LQ iirfilt_crcf iirfilt_crcf_create_prototype(int ftype, int btype, int format, unsigned int order, float fc, float f0, float ap, float as);
LQ int iirfilt_crcf_destroy(iirfilt_crcf q);
/* iirfilt_rrrf (ModemFMStereo.cpp:158-159, 268-269): first-order de-emphasis sections */
typedef struct iirfilt_rrrf_s *iirfilt_rrrf;
LQ iirfilt_rrrf iirfilt_rrrf_create(float *b, unsigned int nb, float *a, unsigned int na);
LQ int iirfilt_rrrf_destroy(iirfilt_rrrf q);
LQ int iirfilt_rrrf_execute(iirfilt_rrrf q, float x, float *y);
/* design entry point, for pinning the product's host-side design (liquid.h:2466) */
LQ int cheby2_azpkf(unsigned int n, float es, lq_cf *z, lq_cf *p, lq_cf *k);
LQ float iirdes_freqprewarp(int btype, float fc, float f0);
LQ int bilinear_zpkf(lq_cf *za, unsigned int nza, lq_cf *pa, unsigned int npa, lq_cf ka, float m, lq_cf *zd, lq_cf *pd, lq_cf *kd);
LQ int iirdes_dzpk_lp2bp(lq_cf *zd, lq_cf *pd, unsigned int n, float f0, lq_cf *zdt, lq_cf *pdt);
LQ int iirdes_dzpk2sosf(lq_cf *zd, lq_cf *pd, unsigned int n, lq_cf kd, float *b, float *a);
LQ int liquid_iirdes(int ftype, int btype, int format, unsigned int n, float fc, float f0, float ap, float as, float *b, float *a);
LQ int iirfilt_crcf_execute(iirfilt_crcf q, lq_cf x, lq_cf *y);
LQ int iirfilt_crcf_execute_block(iirfilt_crcf q, lq_cf *x, unsigned int n, lq_cf *y);

/* firhilbf, liquid.h:3062, 3091 */
typedef struct firhilbf_s *firhilbf;
LQ firhilbf firhilbf_create(unsigned int m, float as);
LQ int firhilbf_destroy(firhilbf q);
LQ int firhilbf_c2r_execute(firhilbf q, lq_cf x, float *y0_lsb, float *y1_usb);
LQ int firhilbf_r2c_execute(firhilbf q, float x, lq_cf *y);             /* ModemFMStereo.cpp:232 */

/* ampmodem, liquid.h:8283-8324 (ModemDSB.cpp:7,38: ampmodem_create(0.5, LIQUID_AMPMODEM_DSB = 0, 1) + ampmodem_demodulate) */
typedef struct ampmodem_s *ampmodem;
LQ ampmodem ampmodem_create(float mod_index, int type, int suppressed_carrier);
LQ int ampmodem_destroy(ampmodem q);
LQ int ampmodem_demodulate(ampmodem q, lq_cf y, float *x);

/* fft, liquid.h:1531, 1582, 1609 */
typedef struct fftplan_s *fftplan;
LQ fftplan fft_create_plan(unsigned int n, lq_cf *x, lq_cf *y, int dir, int flags);
LQ int fft_destroy_plan(fftplan p);
LQ int fft_execute(fftplan p);

/* polyphase filterbank channelizer (SDRPostThread.cpp:401-455: firpfbch_crcf analyzer; LIQUID_ANALYZER = 0) */
typedef struct firpfbch_crcf_s *firpfbch_crcf;
LQ firpfbch_crcf firpfbch_crcf_create_kaiser(int type, unsigned int M, unsigned int m, float As);
LQ int firpfbch_crcf_destroy(firpfbch_crcf q);
LQ int firpfbch_crcf_reset(firpfbch_crcf q);
LQ int firpfbch_crcf_analyzer_execute(firpfbch_crcf q, lq_cf *x, lq_cf *y);
typedef struct firpfbch2_crcf_s *firpfbch2_crcf;
LQ firpfbch2_crcf firpfbch2_crcf_create_kaiser(int type, unsigned int M, unsigned int m, float As);
LQ int firpfbch2_crcf_destroy(firpfbch2_crcf q);
LQ int firpfbch2_crcf_execute(firpfbch2_crcf q, lq_cf *x, lq_cf *y);

#ifdef __cplusplus
}
#endif

#endif
