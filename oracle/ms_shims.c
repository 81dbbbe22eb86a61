/* oracle/ms_shims.c -- TEST INFRASTRUCTURE ONLY (see oracle/build_ref.py).
 *
 * The reference's vendored liquid-dsp 1.5.0 objects were compiled by MinGW and call their libc/libm imports
 * with the Microsoft x64 convention.  build_ref.py renames every import S to ms_S; this file provides ms_S
 * as an ms_abi function that forwards to glibc.  printf-family calls (only reached on liquid error paths)
 * are swallowed.
 */
#define _GNU_SOURCE
#include <complex.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MS __attribute__((ms_abi))

__asm__(".text\n.globl ms____chkstk_ms\n.type ms____chkstk_ms,@function\nms____chkstk_ms:\n\tret\n");

/* stdio: error paths only */
MS int ms___mingw_vfprintf(void *f, const char *fmt, void *ap) { (void)f; (void)fmt; (void)ap; return 0; }
MS int ms___mingw_vfscanf(void *f, const char *fmt, void *ap) { (void)f; (void)fmt; (void)ap; return 0; }
MS int ms___mingw_vsprintf(char *s, const char *fmt, void *ap) { (void)fmt; (void)ap; if (s) s[0] = 0; return 0; }
MS void *ms_fopen(const char *p, const char *m) { return fopen(p, m); }
MS int ms_fclose(void *f) { return fclose((FILE *)f); }
MS int ms_feof(void *f) { return feof((FILE *)f); }
MS size_t ms_fwrite(const void *p, size_t a, size_t b, void *f) { return fwrite(p, a, b, (FILE *)f); }
static MS void *iob_func(unsigned i) { return i == 0 ? (void *)stdin : (i == 1 ? (void *)stdout : (void *)stderr); }
static MS void assert_func(const char *msg, const char *file, unsigned line) {
    fprintf(stderr, "liquid assert: %s (%s:%u)\n", msg ? msg : "?", file ? file : "?", line);
    abort();
}
void *__imp___acrt_iob_func = (void *)iob_func;
void *__imp__assert = (void *)assert_func;

/* memory / strings */
MS void *ms_malloc(size_t n) { return malloc(n); }
MS void *ms_calloc(size_t a, size_t b) { return calloc(a, b); }
MS void *ms_realloc(void *p, size_t n) { return realloc(p, n); }
MS void ms_free(void *p) { free(p); }
MS void *ms_memcpy(void *d, const void *s, size_t n) { return memcpy(d, s, n); }
MS void *ms_memmove(void *d, const void *s, size_t n) { return memmove(d, s, n); }
MS void *ms_memset(void *d, int c, size_t n) { return memset(d, c, n); }
MS int ms_strcmp(const char *a, const char *b) { return strcmp(a, b); }
MS size_t ms_strlen(const char *a) { return strlen(a); }
MS int ms_rand(void) { return rand() & 0x7fff; }
MS div_t ms_div(int a, int b) { return div(a, b); }

typedef MS int (*ms_cmp_t)(const void *, const void *);
static __thread ms_cmp_t g_cmp;
static int cmp_tramp(const void *a, const void *b) { return g_cmp(a, b); }
MS void ms_qsort(void *base, size_t n, size_t sz, ms_cmp_t cmp) { g_cmp = cmp; qsort(base, n, sz, cmp_tramp); }

/* libm, float */
MS float ms_atan2f(float y, float x) { return atan2f(y, x); }
MS float ms_cosf(float x) { return cosf(x); }
MS float ms_sinf(float x) { return sinf(x); }
MS float ms_tanf(float x) { return tanf(x); }
MS float ms_coshf(float x) { return coshf(x); }
MS float ms_sinhf(float x) { return sinhf(x); }
MS float ms_tanhf(float x) { return tanhf(x); }
MS float ms_expf(float x) { return expf(x); }
MS float ms_logf(float x) { return logf(x); }
MS float ms_log10f(float x) { return log10f(x); }
MS float ms_erff(float x) { return erff(x); }
MS float ms_erfcf(float x) { return erfcf(x); }
MS float ms_lgammaf(float x) { return lgammaf(x); }
MS void ms_sincosf(float x, float *s, float *c) { sincosf(x, s, c); }
/* libm, double */
MS double ms_cos(double x) { return cos(x); }
MS double ms_sin(double x) { return sin(x); }
MS double ms_exp(double x) { return exp(x); }
MS double ms_log(double x) { return log(x); }
MS double ms_log10(double x) { return log10(x); }
MS double ms_pow(double x, double y) { return pow(x, y); }
MS double ms_hypot(double x, double y) { return hypot(x, y); }
MS double ms_lgamma(double x) { return lgamma(x); }
/* complex float: 8 bytes by value, same aggregate rule on both sides (gcc ms_abi) */
MS float _Complex ms_ccosf(float _Complex z) { return ccosf(z); }
MS float _Complex ms_csinf(float _Complex z) { return csinf(z); }
MS float _Complex ms_csqrtf(float _Complex z) { return csqrtf(z); }
