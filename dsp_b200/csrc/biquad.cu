// biquad.cu -- K1: fused, time-parallel biquad cascade.
//
// Reference behaviour reproduced (/root/reference): biquad() biquad.h:76-92 (transposed direct
// form II), applied in place per channel by biquad_effect_run[_all] biquad.c:296-315, one pass
// over the block PER STAGE.  Here a cascade of S stages is ONE operator and one read + one write
// of the block.
//
// Per stage (biquad.h:79-81):   r = c0 s + m0;  m0' = m1 + c1 s - c3 r;  m1' = c2 s - c4 r
// The cascade is a linear system with state z = (m0,m1 of every stage) in R^D, D = 2S.  Time is
// cut into chunks of L frames; thread (channel c, chunk j):
//   k_bq_local  runs the cascade over its chunk from ZERO state, keeps only the end state b_j
//   k_bq_scan   one CTA per channel: z_{j+1} = M_c z_j + b_j, z_0 = carried state, M_c = A_c^L, evaluated
//               hierarchically in groups of 8 chunks (A_c = zero-input transition of the cascade, powers
//               M_c^1..M_c^8 taken on the host in long double)
//   k_bq_apply  re-runs the cascade over the chunk from its TRUE start state z_j and stores
// so every output sample is produced by the reference's own recurrence, started from a state
// that differs from the sequential one only by rounding in the scan (|eig A| < 1).
// Lanes run along channels: every global access is a coalesced row of the interleaved block.
#include "common.cuh"
#include "ops.h"

namespace dspb200 {

constexpr int BQ_L = 32;          // frames per chunk
constexpr int BQ_MAX_STAGES = 16; // D <= 32: one lane per state component in the scan

// coef layout: [stage][5][C]; state layouts: b, zin: [c][chunk][D]; zstate: [c][D]
//
// Stages are processed in groups of G, software-pipelined across the samples of the chunk (stage g works
// on sample t-g at step t): the G recurrences are independent dependency chains, so the FP64 pipe sees G-fold
// instruction-level parallelism instead of one serial chain, and the G*5 coefficients of a group are fetched
// together (one exposed load latency per group instead of one per stage).
template <bool APPLY, int G>
__device__ __forceinline__ void bq_stage_group(double (&y)[BQ_L], int nv, const double *__restrict__ coef, int C, int c, int st0, int ng,
                                               const double *z0, double *zend, bool store_end)
{
	double c0[G], c1[G], c2[G], c3[G], c4[G], m0[G], m1[G];
#pragma unroll
	for (int g = 0; g < G; ++g) {
		const bool on = g < ng;
		const double *cf = coef + (long) (st0 + (on ? g : 0)) * 5 * C + c;
		// a missing stage (cascade length not a multiple of G) is the identity section
		c0[g] = on ? cf[0] : 1.0; c1[g] = on ? cf[C] : 0.0; c2[g] = on ? cf[2 * C] : 0.0;
		c3[g] = on ? cf[3 * C] : 0.0; c4[g] = on ? cf[4 * C] : 0.0;
		m0[g] = (APPLY && on) ? z0[2 * (st0 + g)] : 0.0;
		m1[g] = (APPLY && on) ? z0[2 * (st0 + g) + 1] : 0.0;
	}
#pragma unroll
	for (int t = 0; t < BQ_L + G - 1; ++t) {
#pragma unroll
		for (int g = G - 1; g >= 0; --g) {   // later stages first: stage g reads what stage g-1 wrote one step ago
			const int i = t - g;
			if (i >= 0 && i < BQ_L) {
				if (i < nv) {
					const double s = y[i];
					const double r = c0[g] * s + m0[g];
					m0[g] = m1[g] + c1[g] * s - c3[g] * r;
					m1[g] = c2[g] * s - c4[g] * r;
					y[i] = r;
				}
			}
		}
	}
	if (store_end) {
#pragma unroll
		for (int g = 0; g < G; ++g)
			if (g < ng) { zend[2 * (st0 + g)] = m0[g]; zend[2 * (st0 + g) + 1] = m1[g]; }
	}
}

constexpr int BQ_G = 5;

template <bool APPLY>
__global__ void __launch_bounds__(128) k_bq_chunks(const double *in, double *out, const double *__restrict__ coef,
                                                   const double *zin, double *bout, double *zstate,
                                                   int C, int S, long frames, int n_chunks)
{
	const long idx = (long) blockIdx.x * blockDim.x + threadIdx.x;
	const int c = (int) (idx % C);
	const long j = idx / C;
	if (j >= n_chunks) return;
	const int D = 2 * S;
	const long f0 = j * BQ_L;
	const int nv = (int) ((frames - f0 < BQ_L) ? frames - f0 : BQ_L);

	double y[BQ_L];
#pragma unroll
	for (int i = 0; i < BQ_L; ++i) y[i] = (i < nv) ? in[(f0 + i) * C + c] : 0.0;

	const double *z0 = APPLY ? zin + ((long) c * n_chunks + j) * D : nullptr;
	double *zend = APPLY ? zstate + (long) c * D : bout + ((long) c * n_chunks + j) * D;
	const bool store_end = APPLY ? (j == n_chunks - 1) : true;

	for (int st0 = 0; st0 < S; st0 += BQ_G)
		bq_stage_group<APPLY, BQ_G>(y, nv, coef, C, c, st0, (S - st0 < BQ_G) ? S - st0 : BQ_G, z0, zend, store_end);

	if (APPLY) {
#pragma unroll
		for (int i = 0; i < BQ_L; ++i)
			if (i < nv) out[(f0 + i) * C + c] = y[i];
	}
}

// Chunk start states for one channel per CTA.  With P_k = M_c^k (k = 1..8, M_c = A_c^L the one-chunk
// transition) the serial recurrence z_{j+1} = M z_j + b_j over all chunks is cut into groups of 8 chunks:
//   (1) every warp scans ITS group from a zero state: u_{k+1} = P_1 u_k + b_{8G+k}   (7 serial steps, 16 groups in parallel)
//   (2) warp 0 carries the group totals across groups: Z_{G+1} = P_8 Z_G + t_G       (one serial step per group)
//   (3) every warp rebuilds its chunks' true start states: z_{8G+k} = u_k + P_k Z_G  (independent)
// so the serial depth for 128 chunks is 7 + 16 + 1 matrix-vector products instead of 128.  Lane d owns state
// component d (D <= 32); the vectors live in shared memory.
constexpr int BQ_GRP = 8;
constexpr int BQ_SCAN_WARPS = 16;

// sum_e Pt[e][lane] * v[e] + init, Pt = transposed matrix in shared memory (lane-contiguous rows), four partial sums
__device__ __forceinline__ double bq_col_dot(const double *Pt, int D, int lane, const double *v, double init)
{
	double a0 = init, a1 = 0.0, a2 = 0.0, a3 = 0.0;
	int e = 0;
	for (; e + 4 <= D; e += 4) {
		a0 = fma(Pt[e * D + lane], v[e], a0);
		a1 = fma(Pt[(e + 1) * D + lane], v[e + 1], a1);
		a2 = fma(Pt[(e + 2) * D + lane], v[e + 2], a2);
		a3 = fma(Pt[(e + 3) * D + lane], v[e + 3], a3);
	}
	for (; e < D; ++e) a0 = fma(Pt[e * D + lane], v[e], a0);
	return (a0 + a1) + (a2 + a3);
}

__global__ void __launch_bounds__(32 * BQ_SCAN_WARPS) k_bq_scan(const double *__restrict__ Pw, const double *__restrict__ b, double *__restrict__ zin,
                                                                const double *__restrict__ zstate, int C, int D, int n_chunks)
{
	extern __shared__ double dyn[];
	double *Pt = dyn;                                              // [BQ_GRP][D][D], P_k transposed: Pt[k-1][e][d]
	double (*u)[32] = reinterpret_cast<double (*)[32]>(Pt + BQ_GRP * D * D + ((BQ_GRP * D * D) & 1));   // [WARPS*GRP][32]
	double (*tot)[32] = u + BQ_SCAN_WARPS * BQ_GRP;                // [WARPS][32]
	double (*Z)[32] = tot + BQ_SCAN_WARPS;                         // [WARPS+1][32]
	const int c = blockIdx.x;
	const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const double *bc = b + (long) c * n_chunks * D;
	double *zc = zin + (long) c * n_chunks * D;
	const bool on = lane < D;
	{
		const double *P = Pw + (long) c * BQ_GRP * D * D;
		for (int i = threadIdx.x; i < BQ_GRP * D * D; i += blockDim.x) Pt[i] = P[i];
	}
	if (w == 0) Z[0][lane] = on ? zstate[(long) c * D + lane] : 0.0;
	__syncthreads();

	const int super = BQ_SCAN_WARPS * BQ_GRP;           // chunks per pass of the CTA
	for (int base = 0; base < n_chunks; base += super) {
		const int n_here = (n_chunks - base < super) ? n_chunks - base : super;
		const int n_groups = (n_here + BQ_GRP - 1) / BQ_GRP;
		// (1) group-local scans from zero
		if (w < n_groups) {
			double bk[BQ_GRP];
#pragma unroll
			for (int k = 0; k < BQ_GRP; ++k) {
				const int j = w * BQ_GRP + k;
				bk[k] = (on && j < n_here) ? bc[(long) (base + j) * D + lane] : 0.0;
			}
			double uk = 0.0;
#pragma unroll
			for (int k = 0; k < BQ_GRP; ++k) {
				const int j = w * BQ_GRP + k;
				u[j][lane] = uk;
				__syncwarp();
				if (j < n_here && on) uk = bq_col_dot(Pt, D, lane, u[j], bk[k]);
			}
			tot[w][lane] = uk;
		}
		__syncthreads();
		// (2) carry across the groups of this super-block
		if (w == 0) {
			const double *p8 = Pt + (BQ_GRP - 1) * D * D;
			for (int G = 0; G < n_groups; ++G) {
				const double z = on ? bq_col_dot(p8, D, lane, Z[G], tot[G][lane]) : 0.0;
				Z[G + 1][lane] = z;
				__syncwarp();
			}
		}
		__syncthreads();
		// (3) true chunk start states
		if (w < n_groups && on) {
#pragma unroll
			for (int k = 0; k < BQ_GRP; ++k) {
				const int j = w * BQ_GRP + k;
				if (j < n_here) {
					double z = u[j][lane];
					if (k == 0) z += Z[w][lane];
					else z = bq_col_dot(Pt + (k - 1) * D * D, D, lane, Z[w], z);
					zc[(long) (base + j) * D + lane] = z;
				}
			}
		}
		__syncthreads();
		// the next super-block starts where this one ended -- only if it was full (otherwise we are done)
		if (w == 0 && n_here == super) Z[0][lane] = Z[n_groups][lane];
		__syncthreads();
	}
}

struct BiquadOp : Op {
	int S = 0, D = 0;
	std::vector<double> h_coefs;   // [S][C][5] as handed in (kept so neighbouring cascades can be fused)
	double *d_coef = nullptr, *d_M = nullptr, *d_zstate = nullptr, *d_b = nullptr, *d_zin = nullptr;
	long chunk_cap = 0;

	const char *name() const override { return "biquad"; }
	~BiquadOp() override { dev_free(d_coef); dev_free(d_M); dev_free(d_zstate); dev_free(d_b); dev_free(d_zin); }

	void reset(cudaStream_t st) override
	{
		cudaMemsetAsync(d_zstate, 0, (size_t) channels * D * sizeof(double), st);
	}

	long run(long frames, const double *in, double *out, cudaStream_t st) override
	{
		return run_piece(frames, in, out, st);
	}

	long run_piece(long frames, const double *in, double *out, cudaStream_t st)
	{
		if (frames <= 0) return 0;
		const int C = channels;
		const long n_chunks = (frames + BQ_L - 1) / BQ_L;
		const long threads = n_chunks * C;
		ProfScope prof("biquad", st);
		if (n_chunks == 1) {
			// the chunk starts from the carried state itself: zin == zstate (layout [c][1][D])
			LAUNCH(k_bq_chunks<true>, ceil_div(threads, 128), 128, 0, st, in, out, d_coef, d_zstate, nullptr, d_zstate, C, S, frames, 1);
			return frames;
		}
		if (n_chunks > chunk_cap) {
			dev_free(d_b); dev_free(d_zin);
			d_b = dev_alloc<double>((size_t) n_chunks * C * D, false);
			d_zin = dev_alloc<double>((size_t) n_chunks * C * D, false);
			if (!d_b || !d_zin) return -1;
			chunk_cap = n_chunks;
		}
		{
			ProfScope p1("bq_local", st);
			LAUNCH(k_bq_chunks<false>, ceil_div(threads, 128), 128, 0, st, in, nullptr, d_coef, nullptr, d_b, nullptr, C, S, frames, (int) n_chunks);
		}
		{
			ProfScope p2("bq_scan", st);
			const size_t smem = ((size_t) BQ_GRP * D * D + 1 + (size_t) (BQ_SCAN_WARPS * BQ_GRP + 2 * BQ_SCAN_WARPS + 1) * 32) * sizeof(double);
			static std::atomic<int> configured[64];
			int dev = 0;
			cudaGetDevice(&dev);
			if (!configured[dev & 63].load()) {
				const size_t smem_max = ((size_t) BQ_GRP * 4 * BQ_MAX_STAGES * BQ_MAX_STAGES + 1 + (size_t) (BQ_SCAN_WARPS * BQ_GRP + 2 * BQ_SCAN_WARPS + 1) * 32) * sizeof(double);
				CUDA_TRY(cudaFuncSetAttribute(k_bq_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_max), return -1);
				configured[dev & 63].store(1);
			}
			LAUNCH(k_bq_scan, C, 32 * BQ_SCAN_WARPS, smem, st, d_M, d_b, d_zin, d_zstate, C, D, (int) n_chunks);
		}
		{
			ProfScope p3("bq_apply", st);
			LAUNCH(k_bq_chunks<true>, ceil_div(threads, 128), 128, 0, st, in, out, d_coef, d_zin, nullptr, d_zstate, C, S, frames, (int) n_chunks);
		}
		return frames;
	}
};

// zero-input step of the cascade (same arithmetic as biquad.h:79-81 with s = 0 at stage 0)
static void cascade_zero_input_step(int S, const double *cf /*[S][5]*/, long double *z /*[2S]*/)
{
	long double s = 0.0L;
	for (int st = 0; st < S; ++st) {
		const long double c0 = cf[st * 5 + 0], c1 = cf[st * 5 + 1], c2 = cf[st * 5 + 2], c3 = cf[st * 5 + 3], c4 = cf[st * 5 + 4];
		const long double r = c0 * s + z[2 * st];
		z[2 * st] = z[2 * st + 1] + c1 * s - c3 * r;
		z[2 * st + 1] = c2 * s - c4 * r;
		s = r;
	}
}

Op *make_biquad_op(int slab_channels, int fs, int n_stages, const double *coefs)
{
	if (n_stages < 1 || n_stages > BQ_MAX_STAGES) {
		set_error("biquad: %d stages per operator unsupported (1..%d)", n_stages, BQ_MAX_STAGES);
		return nullptr;
	}
	std::unique_ptr<BiquadOp> op(new BiquadOp());
	const int C = slab_channels, S = n_stages, D = 2 * n_stages;
	op->channels = C; op->fs_in = op->fs_out = fs; op->S = S; op->D = D;
	op->h_coefs.assign(coefs, coefs + (size_t) S * C * 5);

	// coefs arrive as [stage][channel][5]; device wants [stage][5][channel]
	std::vector<double> dev_coef((size_t) S * 5 * C), M((size_t) C * BQ_GRP * D * D);
	std::vector<double> cf((size_t) S * 5);
	std::vector<long double> z(D);
	for (int c = 0; c < C; ++c) {
		for (int st = 0; st < S; ++st)
			for (int k = 0; k < 5; ++k) {
				const double v = coefs[((size_t) st * C + c) * 5 + k];
				dev_coef[((size_t) st * 5 + k) * C + c] = v;
				cf[st * 5 + k] = v;
			}
		// column e of P_k = A_c^(k L) e_e, k = 1..BQ_GRP: iterate the zero-input step of the cascade
		for (int e = 0; e < D; ++e) {
			for (int d = 0; d < D; ++d) z[d] = (d == e) ? 1.0L : 0.0L;
			for (int k = 1; k <= BQ_GRP; ++k) {
				for (int i = 0; i < BQ_L; ++i) cascade_zero_input_step(S, cf.data(), z.data());
				for (int d = 0; d < D; ++d) M[(((size_t) c * BQ_GRP + (k - 1)) * D + e) * D + d] = (double) z[d];   // transposed: [e][d]
			}
		}
	}
	op->d_coef = dev_alloc<double>(dev_coef.size(), false);
	op->d_M = dev_alloc<double>(M.size(), false);
	op->d_zstate = dev_alloc<double>((size_t) C * D, true);
	if (!op->d_coef || !op->d_M || !op->d_zstate) return nullptr;
	CUDA_TRY(cudaMemcpy(op->d_coef, dev_coef.data(), dev_coef.size() * sizeof(double), cudaMemcpyHostToDevice), return nullptr);
	CUDA_TRY(cudaMemcpy(op->d_M, M.data(), M.size() * sizeof(double), cudaMemcpyHostToDevice), return nullptr);
	return op.release();
}

// Two adjacent cascades over the same slab become one operator (one pass over the block) when
// the state still fits one lane per component.  Only legal before the first run()/after reset.
Op *fuse_biquad_ops(Op *a, Op *b)
{
	BiquadOp *x = dynamic_cast<BiquadOp *>(a), *y = dynamic_cast<BiquadOp *>(b);
	if (!x || !y || x->channels != y->channels || x->fs_in != y->fs_in || x->S + y->S > BQ_MAX_STAGES) return nullptr;
	std::vector<double> all(x->h_coefs);
	all.insert(all.end(), y->h_coefs.begin(), y->h_coefs.end());
	return make_biquad_op(x->channels, x->fs_in, x->S + y->S, all.data());
}

}  // namespace dspb200
