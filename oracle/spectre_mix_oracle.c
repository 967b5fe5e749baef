/* CPU oracle (plain C, double precision) for the Spectre spectral-mix forward
 * path.  TEST INFRASTRUCTURE ONLY: built by oracle/Makefile into
 * oracle/libspectre_mix_oracle.so and loaded only by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline leg.  The product path
 * (fft_b200/) never links or loads it.
 *
 * It restates the published definition of the transforms the reference calls
 * through torch.fft (third-party: ATen _fft_r2c/_fft_c2r -> MKL DFTI on CPU;
 * no version pinned by the reference, torch 2.11.0 / MKL 2024.2 in this image):
 *
 *   rfft  (spectre.py:506)  X[k] = sum_n x[n] exp(-2 pi i n k / n_fft), k = 0..n_fft/2,
 *                           x zero-padded (N < n_fft) or truncated (N > n_fft)
 *   gate  (spectre.py:542-545)  Y[k,c] = gate[b, c / d_g, k] * X[k,c]
 *   mem   (spectre.py:548-549)  Y[k,c] += memory[k,c]
 *   irfft (spectre.py:551)  y[n] = (1/n_fft) sum_{k=0}^{n_fft-1} Yfull[k] exp(+2 pi i n k / n_fft)
 *                           with Yfull the Hermitian extension of Y; the imaginary
 *                           parts of bin 0 and bin n_fft/2 (n_fft even) are ignored
 *   slice (spectre.py:553)  rows 0..min(N, n_fft)-1
 *
 * Independent of torch/MKL on purpose: an iterative radix-2 FFT in double for
 * power-of-two n_fft, the O(n^2) definition otherwise.  Pinned against the
 * reference-generated fixtures in tests/golden/ by tests/test_oracle.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef struct { double re, im; } cplx;

static int is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

/* in-place iterative radix-2 DIT; sign = -1 forward, +1 inverse (unnormalised) */
static void fft_pow2(cplx *a, int n, int sign, const cplx *w /* w[k] = exp(-2 pi i k / n), k < n/2 */)
{
    for (int i = 1, j = 0; i < n; i++) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { cplx t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1, step = n / len;
        for (int i = 0; i < n; i += len) {
            for (int k = 0; k < half; k++) {
                cplx tw = w[k * step];
                if (sign > 0) tw.im = -tw.im;
                cplx u = a[i + k], v = a[i + k + half];
                cplx t = { v.re * tw.re - v.im * tw.im, v.re * tw.im + v.im * tw.re };
                a[i + k].re = u.re + t.re;        a[i + k].im = u.im + t.im;
                a[i + k + half].re = u.re - t.re; a[i + k + half].im = u.im - t.im;
            }
        }
    }
}

static void dft_naive(const cplx *in, cplx *out, int n, int sign)
{
    for (int k = 0; k < n; k++) {
        double sr = 0.0, si = 0.0;
        for (int t = 0; t < n; t++) {
            /* reduce the product modulo n before scaling to keep the phase exact */
            long long m = ((long long)k * t) % n;
            double ph = sign * 2.0 * M_PI * (double)m / (double)n;
            double c = cos(ph), s = sin(ph);
            sr += in[t].re * c - in[t].im * s;
            si += in[t].re * s + in[t].im * c;
        }
        out[k].re = sr; out[k].im = si;
    }
}

/* v     [B][N][C] float, channel stride 1
 * gate  [B][NG][F_half] interleaved (re, im) float, NG = C / group_width
 * mem   [F_half][C] interleaved (re, im) float, or NULL
 * out   [B][min(N, n_fft)][C] float
 * returns 0 on success, non-zero on bad arguments. */
int spectre_mix_oracle_f64(const float *v, const float *gate, const float *mem, float *out,
                           int B, int N, int n_fft, int C, int group_width)
{
    if (B < 0 || N < 0 || n_fft < 1 || C < 0 || group_width < 1 || (C % group_width) != 0) return 1;
    const int F_half = n_fft / 2 + 1;
    const int NG = C / group_width;
    const int N_in = N < n_fft ? N : n_fft;
    const int N_out = N_in;
    const int pow2 = is_pow2(n_fft);

    cplx *buf = (cplx *)malloc(sizeof(cplx) * (size_t)n_fft);
    cplx *tmp = (cplx *)malloc(sizeof(cplx) * (size_t)n_fft);
    cplx *w = (cplx *)malloc(sizeof(cplx) * (size_t)(n_fft / 2 + 1));
    if (!buf || !tmp || !w) { free(buf); free(tmp); free(w); return 2; }
    for (int k = 0; k < n_fft / 2 + 1; k++) {
        double ph = -2.0 * M_PI * (double)k / (double)n_fft;
        w[k].re = cos(ph); w[k].im = sin(ph);
    }

    for (int b = 0; b < B; b++) {
        for (int c = 0; c < C; c++) {
            const int g = c / group_width;
            /* rfft with zero padding / truncation (spectre.py:506) */
            for (int n = 0; n < n_fft; n++) {
                buf[n].re = n < N_in ? (double)v[((size_t)b * N + n) * C + c] : 0.0;
                buf[n].im = 0.0;
            }
            if (pow2) fft_pow2(buf, n_fft, -1, w);
            else { dft_naive(buf, tmp, n_fft, -1); memcpy(buf, tmp, sizeof(cplx) * (size_t)n_fft); }
            /* gate multiply and memory add on the half spectrum (spectre.py:545, :549) */
            for (int k = 0; k < F_half; k++) {
                const float *gp = gate + (((size_t)b * NG + g) * F_half + k) * 2;
                double gr = gp[0], gi = gp[1];
                double yr = buf[k].re * gr - buf[k].im * gi;
                double yi = buf[k].re * gi + buf[k].im * gr;
                if (mem) {
                    const float *mp = mem + ((size_t)k * C + c) * 2;
                    yr += mp[0]; yi += mp[1];
                }
                tmp[k].re = yr; tmp[k].im = yi;
            }
            /* Hermitian extension as irfft does it (spectre.py:551) */
            tmp[0].im = 0.0;
            if ((n_fft % 2) == 0) tmp[n_fft / 2].im = 0.0;
            for (int k = F_half; k < n_fft; k++) { tmp[k].re = tmp[n_fft - k].re; tmp[k].im = -tmp[n_fft - k].im; }
            if (pow2) { memcpy(buf, tmp, sizeof(cplx) * (size_t)n_fft); fft_pow2(buf, n_fft, +1, w); }
            else dft_naive(tmp, buf, n_fft, +1);
            for (int n = 0; n < N_out; n++)
                out[((size_t)b * N_out + n) * C + c] = (float)(buf[n].re / (double)n_fft);
        }
    }
    free(buf); free(tmp); free(w);
    return 0;
}
