/* oracle/ref_modem_abi.c -- TEST INFRASTRUCTURE ONLY.
 *
 * System V forwarders sv_<name> for the liquid-dsp 1.5.0 entry points the reference's modem sources call (see
 * oracle/modem_abi.h).  Each forwards to the function of that name in the transplanted reference binary
 * (oracle/_ref/libsdr_ref.so), which uses the Microsoft x64 convention (MS).  Objects are opaque pointers; a complex sample
 * is an 8-byte {real, imag} aggregate by value: System V passes it in one SSE register (as it does liquid.h's
 * std::complex<float> / float _Complex), MS in one integer register -- the compiler does the move on each side.
 * Prototypes restate liquid.h (API text: /root/reference/external/liquid-dsp/include/liquid/liquid.h).
 */
#define MS __attribute__((ms_abi))
typedef struct { float real, imag; } cf;
typedef void *obj;
typedef unsigned int uint;

#define FWD0(ret, name, decl, call) MS ret name decl; ret sv_##name decl { return name call; }

FWD0(obj, ampmodem_create, (float mod_index, int type, int suppressed_carrier), (mod_index, type, suppressed_carrier))
FWD0(int, ampmodem_demodulate, (obj q, cf y, float *x), (q, y, x))
FWD0(int, ampmodem_destroy, (obj q), (q))
FWD0(uint, estimate_req_filter_len, (float df, float as), (df, as))
FWD0(obj, firfilt_rrrf_create, (float *h, uint n), (h, n))
FWD0(obj, firfilt_rrrf_create_dc_blocker, (uint m, float as), (m, as))
FWD0(int, firfilt_rrrf_destroy, (obj q), (q))
FWD0(int, firfilt_rrrf_execute, (obj q, float *y), (q, y))
FWD0(int, firfilt_rrrf_push, (obj q, float x), (q, x))
FWD0(int, firhilbf_c2r_execute, (obj q, cf x, float *y0, float *y1), (q, x, y0, y1))
FWD0(obj, firhilbf_create, (uint m, float as), (m, as))
FWD0(int, firhilbf_destroy, (obj q), (q))
FWD0(int, firhilbf_r2c_execute, (obj q, float x, cf *y), (q, x, y))
FWD0(obj, freqdem_create, (float kf), (kf))
FWD0(int, freqdem_demodulate_block, (obj q, cf *r, uint n, float *m), (q, r, n, m))
FWD0(int, freqdem_destroy, (obj q), (q))
FWD0(obj, iirfilt_crcf_create_lowpass, (uint order, float fc), (order, fc))
FWD0(obj, iirfilt_crcf_create_prototype, (int ftype, int btype, int format, uint order, float fc, float f0, float ap, float as),
     (ftype, btype, format, order, fc, f0, ap, as))
FWD0(int, iirfilt_crcf_destroy, (obj q), (q))
FWD0(int, iirfilt_crcf_execute, (obj q, cf x, cf *y), (q, x, y))
FWD0(obj, iirfilt_rrrf_create, (float *b, uint nb, float *a, uint na), (b, nb, a, na))
FWD0(int, iirfilt_rrrf_destroy, (obj q), (q))
FWD0(int, iirfilt_rrrf_execute, (obj q, float x, float *y), (q, x, y))
FWD0(int, liquid_firdes_kaiser, (uint n, float fc, float as, float mu, float *h), (n, fc, as, mu, h))
FWD0(obj, msresamp_cccf_create, (float r, float as), (r, as))
FWD0(int, msresamp_cccf_destroy, (obj q), (q))
FWD0(int, msresamp_cccf_execute, (obj q, cf *x, uint nx, cf *y, uint *ny), (q, x, nx, y, ny))
FWD0(obj, msresamp_rrrf_create, (float r, float as), (r, as))
FWD0(int, msresamp_rrrf_destroy, (obj q), (q))
FWD0(int, msresamp_rrrf_execute, (obj q, float *x, uint nx, float *y, uint *ny), (q, x, nx, y, ny))
FWD0(int, nco_crcf_cexpf, (obj q, cf *y), (q, y))
FWD0(obj, nco_crcf_create, (int type), (type))
FWD0(int, nco_crcf_destroy, (obj q), (q))
FWD0(float, nco_crcf_get_frequency, (obj q), (q))
FWD0(int, nco_crcf_mix_down, (obj q, cf x, cf *y), (q, x, y))
FWD0(int, nco_crcf_mix_up, (obj q, cf x, cf *y), (q, x, y))
FWD0(int, nco_crcf_pll_set_bandwidth, (obj q, float bw), (q, bw))
FWD0(int, nco_crcf_pll_step, (obj q, float dphi), (q, dphi))
FWD0(int, nco_crcf_reset, (obj q), (q))
FWD0(int, nco_crcf_set_frequency, (obj q, float dtheta), (q, dtheta))
FWD0(int, nco_crcf_step, (obj q), (q))

/* the visual processors */
FWD0(obj, fft_create_plan, (uint n, cf *x, cf *y, int dir, int flags), (n, x, y, dir, flags))
FWD0(int, fft_execute, (obj p), (p))
FWD0(int, fft_destroy_plan, (obj p), (p))
FWD0(int, nco_crcf_mix_block_up, (obj q, cf *x, cf *y, uint n), (q, x, y, n))
FWD0(int, nco_crcf_mix_block_down, (obj q, cf *x, cf *y, uint n), (q, x, y, n))
FWD0(obj, msresamp_crcf_create, (float r, float as), (r, as))
FWD0(int, msresamp_crcf_destroy, (obj q), (q))
FWD0(int, msresamp_crcf_execute, (obj q, cf *x, uint nx, cf *y, uint *ny), (q, x, nx, y, ny))

/* the ingress */
FWD0(obj, firpfbch2_crcf_create_kaiser, (int type, uint M, uint m, float as), (type, M, m, as))
FWD0(int, firpfbch2_crcf_destroy, (obj q), (q))
FWD0(int, firpfbch2_crcf_execute, (obj q, cf *x, cf *y), (q, x, y))
FWD0(int, firpfbch_crcf_analyzer_execute, (obj q, cf *x, cf *y), (q, x, y))
FWD0(obj, firpfbch_crcf_create_kaiser, (int type, uint M, uint m, float as), (type, M, m, as))
FWD0(int, firpfbch_crcf_destroy, (obj q), (q))
FWD0(obj, iirfilt_crcf_create_dc_blocker, (float alpha), (alpha))
FWD0(int, iirfilt_crcf_execute_block, (obj q, cf *x, uint n, cf *y), (q, x, n, y))
