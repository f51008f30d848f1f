// pcie_probe.cu -- what the host link of this box does for one 8 MiB block in and one out (mode A of bench.py):
// copy engines with 1-D and 2-D (channel-slab) layouts, one direction and both at once, and SM-driven zero-copy
// (kernels reading/writing mapped pinned host memory).   nvcc -O2 -arch=sm_100a -o pcie_probe pcie_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// rows of `w` doubles (w % 2 == 0) at a pitch of `pitch` doubles: host -> device (compact) and back, 16-byte accesses
__global__ void k_pull(const double2 *__restrict__ host, double2 *__restrict__ dev, long rows, int w2, int pitch2)
{
	const long n = rows * w2;
	for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
		const long r = i / w2;
		const int c = (int) (i - r * w2);
		dev[i] = host[r * pitch2 + c];
	}
}
__global__ void k_push(const double2 *__restrict__ dev, double2 *__restrict__ host, long rows, int w2, int pitch2)
{
	const long n = rows * w2;
	for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long) gridDim.x * blockDim.x) {
		const long r = i / w2;
		const int c = (int) (i - r * w2);
		host[r * pitch2 + c] = dev[i];
	}
}

int main()
{
	const long F = 4096, C = 256;
	const size_t bytes = F * C * 8;
	double *hin, *hout, *hwc, *din, *dout;
	CK(cudaHostAlloc(&hin, bytes, cudaHostAllocPortable | cudaHostAllocMapped));
	CK(cudaHostAlloc(&hout, bytes, cudaHostAllocPortable | cudaHostAllocMapped));
	CK(cudaHostAlloc(&hwc, bytes, cudaHostAllocPortable | cudaHostAllocWriteCombined | cudaHostAllocMapped));
	CK(cudaMalloc(&din, bytes));
	CK(cudaMalloc(&dout, bytes));
	for (size_t i = 0; i < bytes / 8; ++i) { hin[i] = (double) i; hwc[i] = (double) i; }
	cudaStream_t s[8];
	for (auto &x : s) CK(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
	const int reps = 50;
	auto run = [&](const char *name, auto fn, double moved_bytes) {
		for (int i = 0; i < 5; ++i) fn();
		CK(cudaDeviceSynchronize());
		const double t0 = now();
		for (int i = 0; i < reps; ++i) fn();
		CK(cudaDeviceSynchronize());
		const double dt = (now() - t0) / reps;
		printf("%-58s %8.1f us  %7.1f GB/s\n", name, dt * 1e6, moved_bytes / dt / 1e9);
	};
	run("H2D 1-D 8 MiB", [&] { CK(cudaMemcpyAsync(din, hin, bytes, cudaMemcpyHostToDevice, s[0])); CK(cudaStreamSynchronize(s[0])); }, bytes);
	run("H2D 1-D 8 MiB (write-combined source)", [&] { CK(cudaMemcpyAsync(din, hwc, bytes, cudaMemcpyHostToDevice, s[0])); CK(cudaStreamSynchronize(s[0])); }, bytes);
	run("D2H 1-D 8 MiB", [&] { CK(cudaMemcpyAsync(hout, dout, bytes, cudaMemcpyDeviceToHost, s[0])); CK(cudaStreamSynchronize(s[0])); }, bytes);
	run("H2D + D2H 1-D together (2 streams)", [&] {
		CK(cudaMemcpyAsync(din, hin, bytes, cudaMemcpyHostToDevice, s[0]));
		CK(cudaMemcpyAsync(hout, dout, bytes, cudaMemcpyDeviceToHost, s[1]));
		CK(cudaStreamSynchronize(s[0])); CK(cudaStreamSynchronize(s[1])); }, 2.0 * bytes);
	run("H2D then D2H 1-D, one stream (sequential)", [&] {
		CK(cudaMemcpyAsync(din, hin, bytes, cudaMemcpyHostToDevice, s[0]));
		CK(cudaMemcpyAsync(hout, dout, bytes, cudaMemcpyDeviceToHost, s[0]));
		CK(cudaStreamSynchronize(s[0])); }, 2.0 * bytes);
	for (int slabs : { 2, 4, 8 }) {
		const size_t w = C / slabs * 8, pitch = C * 8;
		char name[128];
		snprintf(name, sizeof(name), "H2D 2-D, %d channel slabs (rows of %zu B), one stream each", slabs, w);
		run(name, [&] {
			for (int k = 0; k < slabs; ++k) CK(cudaMemcpy2DAsync((char *) din + k * w * F, w, (char *) hin + k * w, pitch, w, F, cudaMemcpyHostToDevice, s[k]));
			for (int k = 0; k < slabs; ++k) CK(cudaStreamSynchronize(s[k])); }, bytes);
		snprintf(name, sizeof(name), "D2H 2-D, %d channel slabs (rows of %zu B), one stream each", slabs, w);
		run(name, [&] {
			for (int k = 0; k < slabs; ++k) CK(cudaMemcpy2DAsync((char *) hout + k * w, pitch, (char *) dout + k * w * F, w, w, F, cudaMemcpyDeviceToHost, s[k]));
			for (int k = 0; k < slabs; ++k) CK(cudaStreamSynchronize(s[k])); }, bytes);
		snprintf(name, sizeof(name), "pipelined 2-D: slab k H2D -> D2H on its stream, %d slabs", slabs);
		run(name, [&] {
			for (int k = 0; k < slabs; ++k) {
				CK(cudaMemcpy2DAsync((char *) din + k * w * F, w, (char *) hin + k * w, pitch, w, F, cudaMemcpyHostToDevice, s[k]));
				CK(cudaMemcpy2DAsync((char *) hout + k * w, pitch, (char *) din + k * w * F, w, w, F, cudaMemcpyDeviceToHost, s[k]));
			}
			for (int k = 0; k < slabs; ++k) CK(cudaStreamSynchronize(s[k])); }, 2.0 * bytes);
	}
	for (int chunks : { 2, 4, 8 }) {
		const size_t cb = bytes / chunks;
		char name[128];
		snprintf(name, sizeof(name), "pipelined 1-D: %d frame chunks, chunk k H2D -> D2H on its stream", chunks);
		run(name, [&] {
			for (int k = 0; k < chunks; ++k) {
				CK(cudaMemcpyAsync((char *) din + k * cb, (char *) hin + k * cb, cb, cudaMemcpyHostToDevice, s[k]));
				CK(cudaMemcpyAsync((char *) hout + k * cb, (char *) din + k * cb, cb, cudaMemcpyDeviceToHost, s[k]));
			}
			for (int k = 0; k < chunks; ++k) CK(cudaStreamSynchronize(s[k])); }, 2.0 * bytes);
	}
	// SM-driven zero-copy
	double2 *mh_in, *mh_out, *mh_wc;
	CK(cudaHostGetDevicePointer((void **) &mh_in, hin, 0));
	CK(cudaHostGetDevicePointer((void **) &mh_out, hout, 0));
	CK(cudaHostGetDevicePointer((void **) &mh_wc, hwc, 0));
	for (int grid : { 148, 592 }) {
		char name[128];
		snprintf(name, sizeof(name), "kernel pull (mapped host -> HBM), %d CTAs x 256, 16 B loads", grid);
		run(name, [&] { k_pull<<<grid, 256, 0, s[0]>>>(mh_in, (double2 *) din, F, C / 2, C / 2); CK(cudaStreamSynchronize(s[0])); }, bytes);
		snprintf(name, sizeof(name), "kernel pull from write-combined, %d CTAs", grid);
		run(name, [&] { k_pull<<<grid, 256, 0, s[0]>>>(mh_wc, (double2 *) din, F, C / 2, C / 2); CK(cudaStreamSynchronize(s[0])); }, bytes);
		snprintf(name, sizeof(name), "kernel push (HBM -> mapped host), %d CTAs x 256, 16 B stores", grid);
		run(name, [&] { k_push<<<grid, 256, 0, s[0]>>>((double2 *) dout, mh_out, F, C / 2, C / 2); CK(cudaStreamSynchronize(s[0])); }, bytes);
		snprintf(name, sizeof(name), "kernel pull + push together (2 streams), %d CTAs each", grid);
		run(name, [&] {
			k_pull<<<grid, 256, 0, s[0]>>>(mh_in, (double2 *) din, F, C / 2, C / 2);
			k_push<<<grid, 256, 0, s[1]>>>((double2 *) dout, mh_out, F, C / 2, C / 2);
			CK(cudaStreamSynchronize(s[0])); CK(cudaStreamSynchronize(s[1])); }, 2.0 * bytes);
	}
	// one synchronous block in channel slabs: how should the copies of one call be ordered / carried?
	for (int slabs : { 2, 4, 8 }) {
		const size_t w = C / slabs * 8, pitch = C * 8;
		const int w2 = (int) (w / 16), pitch2 = (int) (pitch / 16);
		char name[128];
		snprintf(name, sizeof(name), "breadth-first 2-D: all H2D, then D2H per slab stream, %d slabs", slabs);
		run(name, [&] {
			for (int k = 0; k < slabs; ++k) CK(cudaMemcpy2DAsync((char *) din + k * w * F, w, (char *) hin + k * w, pitch, w, F, cudaMemcpyHostToDevice, s[k]));
			for (int k = 0; k < slabs; ++k) CK(cudaMemcpy2DAsync((char *) hout + k * w, pitch, (char *) din + k * w * F, w, w, F, cudaMemcpyDeviceToHost, s[k]));
			for (int k = 0; k < slabs; ++k) CK(cudaStreamSynchronize(s[k])); }, 2.0 * bytes);
		snprintf(name, sizeof(name), "CE H2D slab k -> push kernel slab k (SM stores to host), %d slabs", slabs);
		run(name, [&] {
			for (int k = 0; k < slabs; ++k) {
				CK(cudaMemcpy2DAsync((char *) din + k * w * F, w, (char *) hin + k * w, pitch, w, F, cudaMemcpyHostToDevice, s[k]));
				k_push<<<148, 256, 0, s[k]>>>((double2 *) ((char *) din + k * w * F), mh_out + k * w2, F, w2, pitch2);
			}
			for (int k = 0; k < slabs; ++k) CK(cudaStreamSynchronize(s[k])); }, 2.0 * bytes);
		snprintf(name, sizeof(name), "pull kernel slab k (SM loads from host) -> CE D2H slab k, %d slabs", slabs);
		run(name, [&] {
			for (int k = 0; k < slabs; ++k) {
				k_pull<<<148, 256, 0, s[k]>>>(mh_in + k * w2, (double2 *) ((char *) din + k * w * F), F, w2, pitch2);
				CK(cudaMemcpy2DAsync((char *) hout + k * w, pitch, (char *) din + k * w * F, w, w, F, cudaMemcpyDeviceToHost, s[k]));
			}
			for (int k = 0; k < slabs; ++k) CK(cudaStreamSynchronize(s[k])); }, 2.0 * bytes);
		snprintf(name, sizeof(name), "pull kernel slab k -> push kernel slab k (no copy engine), %d slabs", slabs);
		run(name, [&] {
			for (int k = 0; k < slabs; ++k) {
				k_pull<<<148, 256, 0, s[k]>>>(mh_in + k * w2, (double2 *) ((char *) din + k * w * F), F, w2, pitch2);
				k_push<<<148, 256, 0, s[k]>>>((double2 *) ((char *) din + k * w * F), mh_out + k * w2, F, w2, pitch2);
			}
			for (int k = 0; k < slabs; ++k) CK(cudaStreamSynchronize(s[k])); }, 2.0 * bytes);
	}
	run("copy-engine H2D + kernel push together", [&] {
		CK(cudaMemcpyAsync(din, hin, bytes, cudaMemcpyHostToDevice, s[0]));
		k_push<<<148, 256, 0, s[1]>>>((double2 *) dout, mh_out, F, C / 2, C / 2);
		CK(cudaStreamSynchronize(s[0])); CK(cudaStreamSynchronize(s[1])); }, 2.0 * bytes);
	return 0;
}
