// common.cuh -- shared plumbing for libdspb200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace dspb200 {

// ---- error reporting (thread-local message behind dspb200_last_error()) ------------------
void set_error(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
const char *get_error();

#define CUDA_TRY(expr, fail_action)                                                        \
	do {                                                                                   \
		cudaError_t err__ = (expr);                                                        \
		if (err__ != cudaSuccess) {                                                        \
			::dspb200::set_error("%s:%d: %s: %s", __FILE__, __LINE__, #expr,               \
			                     cudaGetErrorString(err__));                               \
			fail_action;                                                                   \
		}                                                                                  \
	} while (0)

// ---- launch accounting (dspb200_kernel_launches()) ----------------------------------------
// A rejected launch configuration (grid too large, shared memory the device refuses, ...) is a
// non-sticky error that no later CUDA_TRY would see: LAUNCH records it (message + a per-thread flag)
// and the chain's operator loop turns the flag into a failed call (take_launch_error()).
extern std::atomic<long long> g_kernel_launches;
void note_launch_error(const char *kernel, cudaError_t err);
bool take_launch_error();   // true once after a failed launch on this thread
#define LAUNCH(kernel, grid, block, smem, stream, ...)                                     \
	do {                                                                                   \
		kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                        \
		const cudaError_t lerr__ = cudaPeekAtLastError();                                  \
		if (lerr__ != cudaSuccess) ::dspb200::note_launch_error(#kernel, lerr__);          \
		::dspb200::g_kernel_launches.fetch_add(1, std::memory_order_relaxed);              \
	} while (0)

// ---- per-kernel device timing (dspb200_profile_*): CUDA events on the launching stream ------
extern std::atomic<int> g_profile_on;
void prof_mark(const char *name, cudaStream_t st, bool begin);
struct ProfScope {
	const char *name;
	cudaStream_t st;
	bool on;
	ProfScope(const char *n, cudaStream_t s) : name(n), st(s), on(g_profile_on.load(std::memory_order_relaxed) != 0)
	{
		if (on) prof_mark(name, st, true);
	}
	~ProfScope()
	{
		if (on) prof_mark(name, st, false);
	}
};

// ---- device memory helpers ------------------------------------------------------------------
// Device memory goes through a small cache of freed blocks (api.cu: pool_alloc / pool_free): an operator that is
// re-planned, or a chain that is rebuilt with the same shapes, gets the same blocks back.  Beyond the cudaMalloc time
// this keeps the physical placement a shape has had from the start: the same plan measured up to 2-4x slower on
// memory that had been through a few free / allocate cycles of the driver's allocator (third fresh chain of a
// process: 71-638 us per 2048-frame block where the first took 61-81).  Blocks of at least 1 MiB are kept, up to
// DSP_B200_POOL_MB (default 8192) per process; DSP_B200_POOL_MB=0 turns the cache off.
void *pool_alloc(size_t bytes);
void pool_free(void *p);

template <typename T>
static inline T *dev_alloc(size_t n, bool zero = true)
{
	if (n == 0) n = 1;
	T *p = static_cast<T *>(pool_alloc(n * sizeof(T)));
	if (!p) return nullptr;
	if (zero) {
		// allocation is rare; make the clear visible to every (non-blocking) stream
		cudaMemset(p, 0, n * sizeof(T));
		cudaDeviceSynchronize();
	}
	return p;
}

static inline void dev_free(void *p)
{
	if (p) pool_free(p);
}

static inline int ceil_div(long a, long b) { return (int) ((a + b - 1) / b); }

// ---- one operator instance on one shard (one GPU, one contiguous channel slab) -------------
struct Op {
	int channels = 0;      // channels of this shard's slab
	int fs_in = 0, fs_out = 0;
	bool inplace_ok = true;   // run() accepts in == out
	bool host_mode = false;   // set by the chain: a synchronous host call (one block at a time, copied in and out within the call)
	virtual ~Op() {}
	virtual const char *name() const = 0;
	virtual std::string describe() const { return std::string("{\"op\":\"") + name() + "\"}"; }
	// frames in -> frames out; in/out are device pointers to interleaved [frames][channels]
	virtual long run(long frames, const double *in, double *out, cudaStream_t st) = 0;
	virtual long max_out_frames(long in_frames) const { return in_frames; }
	virtual void reset(cudaStream_t st) = 0;
	// resample.c:163-188: flush what the operator still holds.  `zeros` is scratch for `frames`
	// frames of silence.  Returns frames written to `out`, -1 when dry, -2 on error.
	virtual long drain2(long frames, double *zeros, double *out, cudaStream_t st)
	{
		(void) frames; (void) zeros; (void) out; (void) st;
		return -1;
	}
	// Operators that keep work on streams of their own (K2's look-ahead MACs): make `st` wait for everything
	// enqueued there so far.  dspb200_chain_join(): closes a timed region over ALL of the chain's device work.
	virtual int join(cudaStream_t st)
	{
		(void) st;
		return 0;
	}
	// measurement hook (dspb200_debug_read): operator-specific counters, returns how many were written
	virtual int debug_read(long long *out, int max)
	{
		(void) out; (void) max;
		return 0;
	}
};

// twiddle table exp(-2 pi i t / (2N)), t in [0, 2N), resident on the current device
const double2 *twiddles_2n(int N);
// per-pass butterfly twiddles of the N-point transform (layout: fft.cuh, fft_pass_table_size())
const double2 *twiddles_pass(int N);

}  // namespace dspb200
