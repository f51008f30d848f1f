/*
 * oracle/fftw3_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A small double-precision FFT behind the ten FFTW3 entry points the reference
 * uses (see oracle/fftw3.h for the call sites).  It lets the UNMODIFIED reference
 * sources under /root/reference run here as the parity oracle although FFTW3 is
 * not installed.  Semantics honoured (FFTW3 manual, "Real-data DFTs"):
 *   - r2c: out[k] = sum_j in[j] exp(-2 pi i j k / n), k = 0 .. n/2   (unnormalised)
 *   - c2r: out[j] = sum_k X[k] exp(+2 pi i j k / n) with Hermitian extension,
 *          Im of DC (and of Nyquist for even n) ignored, input may be destroyed
 *   - "new-array" execute on arrays other than the planned ones
 *   - planning never touches the arrays
 *
 * Algorithm: mixed-radix Stockham autosort (radix 4/2/3/5 specialised, any other
 * prime through a generic O(R^2) butterfly), twiddles computed in long double.
 * Even-length real transforms use the half-length complex transform plus the
 * usual split/merge step; odd lengths use a full complex transform.
 * Execution is re-entrant (scratch space is thread-local) because fir_p.c runs
 * its partition groups on worker threads (fir_p.c:105-114).
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "fftw3.h"

typedef struct { double re, im; } cpx;

#define MAX_STAGES 64

struct cfft {
	int n;
	int n_stages;
	int radix[MAX_STAGES];
	cpx *tw[MAX_STAGES];      /* per stage: Ns*(R-1) twiddles, index k*(R-1)+(r-1) */
	cpx *roots[MAX_STAGES];   /* per stage with generic radix: R roots of unity */
};

struct oracle_fft_plan_s {
	int n;            /* real length */
	int kind;         /* 0 = r2c, 1 = c2r */
	int half;         /* n even: complex length n/2, else n */
	struct cfft c;
	cpx *split_tw;    /* n even: exp(-2 pi i k / n), k = 0..n/2 */
	double *in_r; fftw_complex *out_c;   /* planned arrays (r2c) */
	fftw_complex *in_c; double *out_r;   /* planned arrays (c2r) */
};

static void unit_root(long num, long den, cpx *w)
{
	/* exp(-2 pi i num/den), reduced to the first octant for accuracy */
	num %= den;
	if (num < 0) num += den;
	const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double) num / (long double) den;
	w->re = (double) cosl(a);
	w->im = (double) -sinl(a);
}

static int cfft_init(struct cfft *c, int n)
{
	memset(c, 0, sizeof(*c));
	c->n = n;
	int m = n, ns = 1;
	while (m > 1) {
		int r;
		if (m % 4 == 0) r = 4;
		else if (m % 2 == 0) r = 2;
		else if (m % 3 == 0) r = 3;
		else if (m % 5 == 0) r = 5;
		else {
			r = 7;
			while (m % r != 0) {
				r += 2;
				if ((long) r * r > m) { r = m; break; }
			}
		}
		if (c->n_stages >= MAX_STAGES) return 1;
		const int s = c->n_stages++;
		c->radix[s] = r;
		c->tw[s] = malloc(sizeof(cpx) * (size_t) ns * (r - 1));
		if (!c->tw[s]) return 1;
		for (int k = 0; k < ns; ++k)
			for (int q = 1; q < r; ++q)
				unit_root((long) q * k, (long) ns * r, &c->tw[s][(size_t) k * (r - 1) + (q - 1)]);
		if (r != 2 && r != 3 && r != 4 && r != 5) {
			c->roots[s] = malloc(sizeof(cpx) * r);
			if (!c->roots[s]) return 1;
			for (int q = 0; q < r; ++q) unit_root(q, r, &c->roots[s][q]);
		}
		ns *= r;
		m /= r;
	}
	return 0;
}

static void cfft_free(struct cfft *c)
{
	for (int s = 0; s < c->n_stages; ++s) {
		free(c->tw[s]);
		free(c->roots[s]);
	}
}

static inline cpx cmul(cpx a, cpx b)
{
	cpx r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re };
	return r;
}

/* One Stockham pass: src -> dst, sub-transform size so far = ns, radix r. */
static void stockham_pass(const struct cfft *c, int s, int ns, const cpx *src, cpx *dst)
{
	const int n = c->n, r = c->radix[s], m = n / r;
	const cpx *tw = c->tw[s];
	if (r == 2) {
		for (int j = 0; j < m; ++j) {
			const int k = j % ns;
			cpx a = src[j], b = src[j + m];
			if (k) b = cmul(b, tw[k]);
			const int j0 = (j - k) * 2 + k;
			dst[j0].re = a.re + b.re; dst[j0].im = a.im + b.im;
			dst[j0 + ns].re = a.re - b.re; dst[j0 + ns].im = a.im - b.im;
		}
	}
	else if (r == 4) {
		for (int j = 0; j < m; ++j) {
			const int k = j % ns;
			cpx v0 = src[j], v1 = src[j + m], v2 = src[j + 2 * m], v3 = src[j + 3 * m];
			if (k) {
				v1 = cmul(v1, tw[k * 3 + 0]);
				v2 = cmul(v2, tw[k * 3 + 1]);
				v3 = cmul(v3, tw[k * 3 + 2]);
			}
			const cpx a0 = { v0.re + v2.re, v0.im + v2.im };
			const cpx a1 = { v0.re - v2.re, v0.im - v2.im };
			const cpx a2 = { v1.re + v3.re, v1.im + v3.im };
			const cpx a3 = { v1.im - v3.im, -(v1.re - v3.re) };  /* (v1 - v3) * (-i) */
			const int j0 = (j - k) * 4 + k;
			dst[j0].re = a0.re + a2.re; dst[j0].im = a0.im + a2.im;
			dst[j0 + ns].re = a1.re + a3.re; dst[j0 + ns].im = a1.im + a3.im;
			dst[j0 + 2 * ns].re = a0.re - a2.re; dst[j0 + 2 * ns].im = a0.im - a2.im;
			dst[j0 + 3 * ns].re = a1.re - a3.re; dst[j0 + 3 * ns].im = a1.im - a3.im;
		}
	}
	else if (r == 3) {
		const double c1 = -0.5, s1 = -0.86602540378443864676372317075294;  /* exp(-2 pi i/3) */
		for (int j = 0; j < m; ++j) {
			const int k = j % ns;
			cpx v0 = src[j], v1 = src[j + m], v2 = src[j + 2 * m];
			if (k) {
				v1 = cmul(v1, tw[k * 2 + 0]);
				v2 = cmul(v2, tw[k * 2 + 1]);
			}
			const cpx t = { v1.re + v2.re, v1.im + v2.im };
			const cpx d = { v1.re - v2.re, v1.im - v2.im };
			const cpx u = { v0.re + c1 * t.re, v0.im + c1 * t.im };
			const cpx w = { -s1 * d.im, s1 * d.re };  /* i*s1*d */
			const int j0 = (j - k) * 3 + k;
			dst[j0].re = v0.re + t.re; dst[j0].im = v0.im + t.im;
			dst[j0 + ns].re = u.re + w.re; dst[j0 + ns].im = u.im + w.im;
			dst[j0 + 2 * ns].re = u.re - w.re; dst[j0 + 2 * ns].im = u.im - w.im;
		}
	}
	else if (r == 5) {
		/* exp(-2 pi i q/5): cos and -sin */
		const double c1 = 0.30901699437494742410229341718282, s1 = -0.95105651629515357211643933337938;
		const double c2 = -0.80901699437494742410229341718282, s2 = -0.58778525229247312916870595463907;
		for (int j = 0; j < m; ++j) {
			const int k = j % ns;
			cpx v0 = src[j], v1 = src[j + m], v2 = src[j + 2 * m], v3 = src[j + 3 * m], v4 = src[j + 4 * m];
			if (k) {
				v1 = cmul(v1, tw[k * 4 + 0]);
				v2 = cmul(v2, tw[k * 4 + 1]);
				v3 = cmul(v3, tw[k * 4 + 2]);
				v4 = cmul(v4, tw[k * 4 + 3]);
			}
			const cpx t1 = { v1.re + v4.re, v1.im + v4.im }, d1 = { v1.re - v4.re, v1.im - v4.im };
			const cpx t2 = { v2.re + v3.re, v2.im + v3.im }, d2 = { v2.re - v3.re, v2.im - v3.im };
			const cpx u1 = { v0.re + c1 * t1.re + c2 * t2.re, v0.im + c1 * t1.im + c2 * t2.im };
			const cpx u2 = { v0.re + c2 * t1.re + c1 * t2.re, v0.im + c2 * t1.im + c1 * t2.im };
			/* i*(s1*d1 + s2*d2) and i*(s2*d1 - s1*d2) */
			const cpx e1 = { s1 * d1.re + s2 * d2.re, s1 * d1.im + s2 * d2.im };
			const cpx e2 = { s2 * d1.re - s1 * d2.re, s2 * d1.im - s1 * d2.im };
			const cpx w1 = { -e1.im, e1.re }, w2 = { -e2.im, e2.re };
			const int j0 = (j - k) * 5 + k;
			dst[j0].re = v0.re + t1.re + t2.re; dst[j0].im = v0.im + t1.im + t2.im;
			dst[j0 + ns].re = u1.re + w1.re; dst[j0 + ns].im = u1.im + w1.im;
			dst[j0 + 4 * ns].re = u1.re - w1.re; dst[j0 + 4 * ns].im = u1.im - w1.im;
			dst[j0 + 2 * ns].re = u2.re + w2.re; dst[j0 + 2 * ns].im = u2.im + w2.im;
			dst[j0 + 3 * ns].re = u2.re - w2.re; dst[j0 + 3 * ns].im = u2.im - w2.im;
		}
	}
	else {
		const cpx *roots = c->roots[s];
		cpx *v = malloc(sizeof(cpx) * r);
		for (int j = 0; j < m; ++j) {
			const int k = j % ns;
			v[0] = src[j];
			for (int q = 1; q < r; ++q) {
				v[q] = src[j + (size_t) q * m];
				if (k) v[q] = cmul(v[q], tw[(size_t) k * (r - 1) + (q - 1)]);
			}
			const int j0 = (j - k) * r + k;
			for (int p = 0; p < r; ++p) {
				double ar = v[0].re, ai = v[0].im;
				int idx = 0;
				for (int q = 1; q < r; ++q) {
					idx += p;
					if (idx >= r) idx -= r;
					ar += v[q].re * roots[idx].re - v[q].im * roots[idx].im;
					ai += v[q].re * roots[idx].im + v[q].im * roots[idx].re;
				}
				dst[j0 + (size_t) p * ns].re = ar;
				dst[j0 + (size_t) p * ns].im = ai;
			}
		}
		free(v);
	}
}

/* Forward complex DFT of c->n points. `a` holds the input; the result is returned in
 * either `a` or `b` (pointer returned). */
static cpx * cfft_forward(const struct cfft *c, cpx *a, cpx *b)
{
	int ns = 1;
	cpx *src = a, *dst = b;
	for (int s = 0; s < c->n_stages; ++s) {
		stockham_pass(c, s, ns, src, dst);
		ns *= c->radix[s];
		cpx *t = src; src = dst; dst = t;
	}
	return src;
}

static __thread cpx *scratch = NULL;
static __thread size_t scratch_len = 0;

static cpx * get_scratch(size_t len)
{
	if (scratch_len < len) {
		free(scratch);
		scratch = malloc(sizeof(cpx) * len);
		scratch_len = (scratch) ? len : 0;
	}
	return scratch;
}

void * fftw_malloc(size_t n)
{
	void *p = NULL;
	if (n == 0) n = 64;
	if (posix_memalign(&p, 64, n) != 0) return NULL;
	memset(p, 0, n);
	return p;
}

void fftw_free(void *p)
{
	free(p);
}

static fftw_plan plan_new(int n, int kind)
{
	if (n < 1) return NULL;
	fftw_plan p = calloc(1, sizeof(*p));
	if (!p) return NULL;
	p->n = n;
	p->kind = kind;
	p->half = (n % 2 == 0) ? n / 2 : n;
	if (cfft_init(&p->c, p->half)) goto fail;
	if (n % 2 == 0) {
		p->split_tw = malloc(sizeof(cpx) * (n / 2 + 1));
		if (!p->split_tw) goto fail;
		for (int k = 0; k <= n / 2; ++k) unit_root(k, n, &p->split_tw[k]);
	}
	return p;
	fail:
	fftw_destroy_plan(p);
	return NULL;
}

fftw_plan fftw_plan_dft_r2c_1d(int n, double *in, fftw_complex *out, unsigned flags)
{
	(void) flags;
	fftw_plan p = plan_new(n, 0);
	if (p) { p->in_r = in; p->out_c = out; }
	return p;
}

fftw_plan fftw_plan_dft_c2r_1d(int n, fftw_complex *in, double *out, unsigned flags)
{
	(void) flags;
	fftw_plan p = plan_new(n, 1);
	if (p) { p->in_c = in; p->out_r = out; }
	return p;
}

void fftw_destroy_plan(fftw_plan p)
{
	if (!p) return;
	cfft_free(&p->c);
	free(p->split_tw);
	free(p);
}

void fftw_execute_dft_r2c(const fftw_plan p, double *in, fftw_complex *out_)
{
	cpx *out = (cpx *) out_;
	const int n = p->n, h = p->half;
	cpx *a = get_scratch((size_t) 2 * h), *b = a + h;
	if (n % 2 == 0) {
		for (int j = 0; j < h; ++j) { a[j].re = in[2 * j]; a[j].im = in[2 * j + 1]; }
		const cpx *z = cfft_forward(&p->c, a, b);
		/* X[k] = E[k] + w^k O[k], E = (Z[k] + conj Z[h-k])/2, O = -i (Z[k] - conj Z[h-k])/2 */
		out[0].re = z[0].re + z[0].im; out[0].im = 0.0;
		out[h].re = z[0].re - z[0].im; out[h].im = 0.0;
		for (int k = 1; k < h; ++k) {
			const cpx zk = z[k], zn = z[h - k];
			const cpx e = { 0.5 * (zk.re + zn.re), 0.5 * (zk.im - zn.im) };
			const cpx o = { 0.5 * (zk.im + zn.im), -0.5 * (zk.re - zn.re) };
			const cpx wo = cmul(p->split_tw[k], o);
			out[k].re = e.re + wo.re;
			out[k].im = e.im + wo.im;
		}
	}
	else {
		for (int j = 0; j < n; ++j) { a[j].re = in[j]; a[j].im = 0.0; }
		const cpx *z = cfft_forward(&p->c, a, b);
		for (int k = 0; k <= n / 2; ++k) out[k] = z[k];
	}
}

void fftw_execute_dft_c2r(const fftw_plan p, fftw_complex *in_, double *out)
{
	const cpx *in = (const cpx *) in_;
	const int n = p->n, h = p->half;
	cpx *a = get_scratch((size_t) 2 * h), *b = a + h;
	if (n % 2 == 0) {
		/* Z'[k] = (X[k] + conj X[h-k]) + i conj(w^k) (X[k] - conj X[h-k]); inverse via conj(DFT(conj(.))) */
		a[0].re = in[0].re + in[h].re;
		a[0].im = -(in[0].re - in[h].re);
		for (int k = 1; k < h; ++k) {
			const cpx xk = in[k], xn = in[h - k];
			const cpx s = { xk.re + xn.re, xk.im - xn.im };
			const cpx d = { xk.re - xn.re, xk.im + xn.im };
			const cpx wc = { p->split_tw[k].re, -p->split_tw[k].im };
			const cpx t = cmul(wc, d);  /* conj(w^k) d */
			/* z = s + i t ; store conj(z) */
			a[k].re = s.re - t.im;
			a[k].im = -(s.im + t.re);
		}
		const cpx *z = cfft_forward(&p->c, a, b);
		for (int j = 0; j < h; ++j) { out[2 * j] = z[j].re; out[2 * j + 1] = -z[j].im; }
	}
	else {
		a[0].re = in[0].re; a[0].im = 0.0;
		for (int k = 1; k <= n / 2; ++k) {
			a[k].re = in[k].re; a[k].im = -in[k].im;           /* conj(X[k]) */
			a[n - k].re = in[k].re; a[n - k].im = in[k].im;    /* conj(conj(X[k])) */
		}
		const cpx *z = cfft_forward(&p->c, a, b);
		for (int j = 0; j < n; ++j) out[j] = z[j].re;
	}
}

void fftw_execute(const fftw_plan p)
{
	if (p->kind == 0) fftw_execute_dft_r2c(p, p->in_r, p->out_c);
	else fftw_execute_dft_c2r(p, p->in_c, p->out_r);
}

int fftw_import_wisdom_from_filename(const char *filename)
{
	(void) filename;
	return 0;
}

int fftw_export_wisdom_to_filename(const char *filename)
{
	(void) filename;
	return 0;
}
