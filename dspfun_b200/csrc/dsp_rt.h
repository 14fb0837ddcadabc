// dsp_rt.h -- thin runtime layer under the planner: device memory, copies, streams.
// Product build: CUDA runtime.  -DDSP_EMULATE (tests/emu only): plain host memory, so the planner, the table
// builders and the C ABI can be exercised on a machine without a GPU.
#pragma once
#include <stddef.h>
#include <string>

#if defined(__CUDACC__) && !defined(DSP_EMULATE)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
namespace dsp {
typedef cudaStream_t rt_stream;
inline bool rt_ok(cudaError_t e, std::string &err, const char *what) {
	if (e == cudaSuccess) return true;
	err = std::string(what) + ": " + cudaGetErrorString(e);
	return false;
}
inline bool rt_init(std::string &err) {
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n < 1) {
		err = std::string("no usable CUDA device (libdspdct has no CPU path): ") + cudaGetErrorString(e);
		cudaGetLastError();
		return false;
	}
	return true;
}
inline int rt_device() { int d = 0; cudaGetDevice(&d); return d; }
inline bool rt_malloc(void **p, size_t bytes, std::string &err) {
	if (getenv("DSP_DCT_TRACE")) { fprintf(stderr, "[dsp_dct] cudaMalloc %zu\n", bytes); fflush(stderr); }
	cudaError_t e = cudaMalloc(p, bytes ? bytes : 1);
	if (getenv("DSP_DCT_TRACE")) { fprintf(stderr, "[dsp_dct] cudaMalloc -> %d\n", (int)e); fflush(stderr); }
	return rt_ok(e, err, "cudaMalloc");
}
inline void rt_free(void *p) { if (p) cudaFree(p); }
inline bool rt_h2d(void *d, const void *h, size_t n, rt_stream s, std::string &err) { return rt_ok(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s), err, "H2D copy"); }
inline bool rt_d2h(void *h, const void *d, size_t n, rt_stream s, std::string &err) { return rt_ok(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s), err, "D2H copy"); }
inline bool rt_d2d(void *dst, const void *src, size_t n, rt_stream s, std::string &err) { return rt_ok(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, s), err, "D2D copy"); }
inline bool rt_zero(void *d, size_t n, rt_stream s, std::string &err) { return rt_ok(cudaMemsetAsync(d, 0, n, s), err, "memset"); }
inline bool rt_sync(rt_stream s, std::string &err) { return rt_ok(cudaStreamSynchronize(s), err, "stream sync"); }
inline bool rt_is_device_ptr(const void *p) {
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
	return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
inline void *rt_host_alloc(size_t bytes) { void *p = nullptr; if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; } return p; }
inline void rt_host_free(void *p) { if (p) cudaFreeHost(p); }
}  // namespace dsp
#else
#include <stdlib.h>
#include <string.h>
namespace dsp {
typedef void *rt_stream;
inline bool rt_init(std::string &) { return true; }
inline int rt_device() { return 0; }
inline bool rt_malloc(void **p, size_t bytes, std::string &err) { *p = malloc(bytes ? bytes : 1); if (!*p) { err = "malloc failed"; return false; } return true; }
inline void rt_free(void *p) { free(p); }
inline bool rt_h2d(void *d, const void *h, size_t n, rt_stream, std::string &) { memcpy(d, h, n); return true; }
inline bool rt_d2h(void *h, const void *d, size_t n, rt_stream, std::string &) { memcpy(h, d, n); return true; }
inline bool rt_d2d(void *dst, const void *src, size_t n, rt_stream, std::string &) { memmove(dst, src, n); return true; }
inline bool rt_zero(void *d, size_t n, rt_stream, std::string &) { memset(d, 0, n); return true; }
inline bool rt_sync(rt_stream, std::string &) { return true; }
inline bool rt_is_device_ptr(const void *) { return false; }
inline void *rt_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
inline void rt_host_free(void *p) { free(p); }
}  // namespace dsp
#endif
