// dct_tma.cuh -- the async-copy plumbing shared by the persistent ring kernels (dct_ring.cuh, dct_colring.cuh):
// mbarrier, 1-D bulk copies (cp.async.bulk) and tiled tensor copies (cp.async.bulk.tensor, TMA) between global and
// shared memory.  GPU: inline PTX for sm_100a, tensor maps encoded by the driver (cuTensorMapEncodeTiled through
// cudaGetDriverEntryPoint: no link-time dependency on libcuda).  Emulation (tests/emu): the same calls as plain loops,
// so that the kernels' box geometry and index arithmetic are exercised on a machine without a GPU.
#pragma once
#include "dct_core.cuh"
#include <string.h>
#include <string>
#if DSP_GPU
#include <cuda.h>
#else
#include <vector>
#endif

namespace dsp {

// ------------------------------------------------------------------------------------------------ mbarrier, 1-D bulk copy
#if DSP_GPU
DSP_DEV uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
DSP_DEV void mbar_init(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
DSP_DEV void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
DSP_DEV void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DSP_DEV void mbar_wait(uint64_t *bar, uint32_t parity) {
	uint32_t ok;
	do {      // try_wait suspends the thread in hardware for a bounded time; no labels, so the block may be duplicated freely
		asm volatile(
		    "{\n"
		    ".reg .pred p;\n"
		    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
		    "selp.u32 %0, 1, 0, p;\n"
		    "}\n"
		    : "=r"(ok)
		    : "r"(smem_u32(bar)), "r"(parity)
		    : "memory");
	} while (!ok);
}
// global -> shared bulk copy (bytes % 16 == 0, both addresses 16-byte aligned); completes on `bar`
DSP_DEV void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
	             "l"(src), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}
// orders this thread's (and, after a barrier, its group's) generic-proxy accesses to shared memory before later
// async-proxy (bulk copy) writes to the same locations
DSP_DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif


// ------------------------------------------------------------------------------------------------ tensor maps
// A tiled view of a float array: rank <= 4 dimensions (dimension 0 contiguous), byte strides for dimensions 1.., and
// the box (tile) one copy moves.  Elements of a box outside the array read as zero and are not written.
struct TmaView {
	void *base;
	int rank;
	unsigned long long dims[4], strides[4];     // strides[0] is implied (4 bytes)
	unsigned box[4];
};
#if DSP_GPU
typedef CUtensorMap TmaDesc;
// swizzle128: the box lands with the 128-byte swizzle (16-byte chunk index ^= row % 8; box rows of exactly 128 bytes,
// destination 1024-byte aligned), the layout the tensor core's SWIZZLE_128B operand descriptors read
inline bool tma_encode(TmaDesc *d, const TmaView &v, std::string &err, bool swizzle128 = false) {
	typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
	                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
	                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
	static EncodeFn fn = nullptr;
	if (!fn) {
		void *p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) {
			err = "cuTensorMapEncodeTiled is not available from this driver";
			return false;
		}
		fn = (EncodeFn)p;
	}
	cuuint64_t dims[4], strides[3];
	cuuint32_t box[4], es[4];
	for (int i = 0; i < v.rank; i++) { dims[i] = v.dims[i]; box[i] = v.box[i]; es[i] = 1; if (i) strides[i - 1] = v.strides[i]; }
	const CUresult rc = fn(d, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)v.rank, v.base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
	                       swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (rc != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)rc) + ")"; return false; }
	return true;
}
// box at coordinates (c0, c1, c2[, c3]) -> shared memory (128-byte aligned); completes on `bar` with the box's full byte count
DSP_DEV void tma_load2(void *dst, const TmaDesc *map, int c0, int c1, uint64_t *bar) {
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
	             "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
	             : "memory");
}
DSP_DEV void tma_load3(void *dst, const TmaDesc *map, int c0, int c1, int c2, uint64_t *bar) {
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
	             "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
	             : "memory");
}
DSP_DEV void tma_load4(void *dst, const TmaDesc *map, int c0, int c1, int c2, int c3, uint64_t *bar) {
	asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(dst)),
	             "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
	             : "memory");
}
// shared memory -> box at the coordinates; part of the thread's current bulk group
DSP_DEV void tma_store3(const TmaDesc *map, const void *src, int c0, int c1, int c2) {
	asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1),
	             "r"(c2)
	             : "memory");
}
DSP_DEV void tma_store4(const TmaDesc *map, const void *src, int c0, int c1, int c2, int c3) {
	asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(c0),
	             "r"(c1), "r"(c2), "r"(c3)
	             : "memory");
}
// L2 eviction-priority policies for the cache-hinted copies below
DSP_DEV uint64_t l2_policy_evict_last() {
	uint64_t p;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
	return p;
}
DSP_DEV uint64_t l2_policy_evict_first() {
	uint64_t p;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
	return p;
}
DSP_DEV void tma_load4_hint(void *dst, const TmaDesc *map, int c0, int c1, int c2, int c3, uint64_t *bar, uint64_t policy) {
	asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5}], [%6], %7;" ::"r"(smem_u32(dst)),
	             "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)), "l"(policy)
	             : "memory");
}
DSP_DEV void tma_store4_hint(const TmaDesc *map, const void *src, int c0, int c1, int c2, int c3, uint64_t policy) {
	asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%2, %3, %4, %5}], [%1], %6;" ::"l"(map), "r"(smem_u32(src)), "r"(c0),
	             "r"(c1), "r"(c2), "r"(c3), "l"(policy)
	             : "memory");
}
// the 128-byte line at p (aligned) will not be read again before it is rewritten: the L2 may drop it without write-back
DSP_DEV void discard_l2(const void *p) { asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory"); }
DSP_DEV void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
DSP_DEV void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }    // sources of all groups read
DSP_DEV void tma_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }          // all groups complete
#else
typedef TmaView TmaDesc;
inline bool tma_encode(TmaDesc *d, const TmaView &v, std::string &, bool = false) { *d = v; return true; }
inline void tma_emu_copy(const TmaDesc *m, float *smem, const int *c, bool load) {
	unsigned long long st[4] = {4, 0, 0, 0};
	int cc[4] = {0, 0, 0, 0};
	unsigned bx[4] = {1, 1, 1, 1};
	unsigned long long dm[4] = {1, 1, 1, 1};
	for (int i = 0; i < m->rank; i++) { cc[i] = c[i]; bx[i] = m->box[i]; dm[i] = m->dims[i]; if (i) st[i] = m->strides[i]; }
	size_t s = 0;
	for (unsigned i3 = 0; i3 < bx[3]; i3++)
		for (unsigned i2 = 0; i2 < bx[2]; i2++)
			for (unsigned i1 = 0; i1 < bx[1]; i1++)
				for (unsigned i0 = 0; i0 < bx[0]; i0++, s++) {
					const long long x0 = cc[0] + (long long)i0, x1 = cc[1] + (long long)i1, x2 = cc[2] + (long long)i2, x3 = cc[3] + (long long)i3;
					const bool in = x0 >= 0 && x1 >= 0 && x2 >= 0 && x3 >= 0 && (unsigned long long)x0 < dm[0] && (unsigned long long)x1 < dm[1] &&
					                (unsigned long long)x2 < dm[2] && (unsigned long long)x3 < dm[3];
					float *g = (float *)((char *)m->base + x0 * 4 + x1 * (long long)st[1] + x2 * (long long)st[2] + x3 * (long long)st[3]);
					if (load) smem[s] = in ? *g : 0.0f;
					else if (in) *g = smem[s];
				}
}
inline void tma_load3(void *dst, const TmaDesc *map, int c0, int c1, int c2, void *) { const int c[4] = {c0, c1, c2, 0}; tma_emu_copy(map, (float *)dst, c, true); }
inline void tma_load4(void *dst, const TmaDesc *map, int c0, int c1, int c2, int c3, void *) { const int c[4] = {c0, c1, c2, c3}; tma_emu_copy(map, (float *)dst, c, true); }
inline void tma_store3(const TmaDesc *map, const void *src, int c0, int c1, int c2) { const int c[4] = {c0, c1, c2, 0}; tma_emu_copy(map, (float *)src, c, false); }
inline void tma_store4(const TmaDesc *map, const void *src, int c0, int c1, int c2, int c3) { const int c[4] = {c0, c1, c2, c3}; tma_emu_copy(map, (float *)src, c, false); }
inline unsigned long long l2_policy_evict_last() { return 0; }
inline unsigned long long l2_policy_evict_first() { return 0; }
inline void tma_load4_hint(void *dst, const TmaDesc *map, int c0, int c1, int c2, int c3, void *bar, unsigned long long) { tma_load4(dst, map, c0, c1, c2, c3, bar); }
inline void tma_store4_hint(const TmaDesc *map, const void *src, int c0, int c1, int c2, int c3, unsigned long long) { tma_store4(map, src, c0, c1, c2, c3); }
#endif

// Every thread that read or wrote a shared-memory buffer through the generic proxy runs this BEFORE the barrier after
// which one thread hands the buffer to the copy engine (async proxy), for a refill or for a store: the barrier orders
// the threads, the fence orders each thread's own accesses against the other proxy.  (A fence by the issuing thread
// alone is not enough: measured as nondeterministic results of the column ring once the timing got tight.)
#if DSP_GPU
#define RING_PROXY_FENCE() fence_proxy_async()
#else
#define RING_PROXY_FENCE() ((void)0)
#endif

}  // namespace dsp
