// oracle/ocl_host.cpp -- TEST INFRASTRUCTURE. Runs the reference's OWN OpenCL C kernels (bin/ocl_kernels.cl, the ETC1S GPU seam,
// SURVEY.md 2.1) on the host, so the CUDA implementations of the same five entry points have a "reference"-kind oracle even
// though no OpenCL runtime exists in this image.
//
// oracle/Makefile generates oracle/_ref/ocl_kernels_host.inc from the .cl where it lies under /root/reference with one
// mechanical rewrite -- OpenCL vector literals `(color_rgba)(a,b,c,d)` become `make_color_rgba(a,b,c,d)` -- and this file
// supplies the handful of OpenCL C built-ins the kernels use (uchar4, float3, min/max/clamp, address-space qualifiers,
// get_global_id). Nothing of the kernels' logic is restated here. Compiled with -ffp-contract=off (OpenCL leaves contraction
// to the implementation; the CUDA path is built without FMA contraction, so that is the variant pinned).
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>

namespace oclref
{
	typedef unsigned char uchar;
	typedef unsigned short ushort;
	typedef unsigned int uint;
	typedef unsigned long ulong;

	struct uchar4
	{
		uchar x, y, z, w;
		uchar4() {}
		uchar4(int v) : x((uchar)v), y((uchar)v), z((uchar)v), w((uchar)v) {} // scalar broadcast, e.g. `color_rgba min_color = 255;`
	};
	static inline uchar4 make_color_rgba(int a, int b, int c, int d) { uchar4 r; r.x = (uchar)a; r.y = (uchar)b; r.z = (uchar)c; r.w = (uchar)d; return r; }
	struct float3 { float x, y, z; };

	template <typename T> static inline T min(T a, T b) { return (b < a) ? b : a; }
	template <typename T> static inline T max(T a, T b) { return (a < b) ? b : a; }
	template <typename T> static inline T clamp(T v, T lo, T hi) { return min(max(v, lo), hi); }
	static inline uchar4 min(uchar4 a, uchar4 b) { return make_color_rgba(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z), min(a.w, b.w)); }
	static inline uchar4 max(uchar4 a, uchar4 b) { return make_color_rgba(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z), max(a.w, b.w)); }

	static thread_local unsigned int g_global_id = 0;
	static inline unsigned int get_global_id(int) { return g_global_id; }

#define kernel
#define global
#define constant const
#include "_ref/ocl_kernels_host.inc"
#undef kernel
#undef global
#undef constant
}

#define OCL_API extern "C" __attribute__((visibility("default")))

template <typename F> static void for_each_item(uint32_t n, uint32_t threads, F f)
{
	if (threads <= 1) { for (uint32_t i = 0; i < n; i++) { oclref::g_global_id = i; f(); } return; }
	std::atomic<uint32_t> next(0);
	std::vector<std::thread> pool;
	for (uint32_t t = 0; t < threads; t++)
		pool.emplace_back([&]() {
			for (;;)
			{
				const uint32_t first = next.fetch_add(256);
				if (first >= n) break;
				const uint32_t last = first + 256 < n ? first + 256 : n;
				for (uint32_t i = first; i < last; i++) { oclref::g_global_id = i; f(); }
			}
		});
	for (auto& th : pool) th.join();
}

// opencl_encode_etc1s_blocks (encoder/basisu_opencl.cpp:932) -> kernel encode_etc1s_blocks (ocl_kernels.cl:984)
OCL_API void oclref_encode_etc1s_blocks(const void* pBlocks, uint32_t n, void* pOut, int perceptual, int total_perms, uint32_t threads)
{
	oclref::encode_etc1s_param_struct p; p.m_total_blocks = n; p.m_perceptual = perceptual; p.m_total_perms = total_perms;
	memset(pOut, 0, (size_t)n * 8);
	for_each_item(n, threads, [&]() { oclref::encode_etc1s_blocks(&p, (const oclref::pixel_block*)pBlocks, (oclref::etc_block*)pOut); });
}

// opencl_encode_etc1s_pixel_clusters (opencl.cpp:979) -> kernel encode_etc1s_from_pixel_cluster (cl:1013).
// The kernel leaves the selector bytes (4..7) of each output block unwritten; callers compare bytes 0..3 only.
OCL_API void oclref_encode_etc1s_pixel_clusters(const void* pClusters, uint32_t n, const void* pPixels, const uint32_t* pWeights, void* pOut, int perceptual, int total_perms, uint32_t threads)
{
	oclref::encode_etc1s_param_struct p; p.m_total_blocks = n; p.m_perceptual = perceptual; p.m_total_perms = total_perms;
	memset(pOut, 0, (size_t)n * 8);
	for_each_item(n, threads, [&]() { oclref::encode_etc1s_from_pixel_cluster(&p, (const oclref::pixel_cluster*)pClusters, (const oclref::color_rgba*)pPixels, pWeights, (oclref::etc_block*)pOut); });
}

// opencl_refine_endpoint_clusterization (opencl.cpp:1051) -> kernel refine_endpoint_clusterization (cl:1063)
OCL_API void oclref_refine_endpoint_clusterization(const void* pBlocks, uint32_t n, const void* pBlock_info, const void* pClusters, const uint32_t* pSorted_block_indices, uint32_t* pOut, int perceptual, uint32_t threads)
{
	oclref::rec_param_struct p; p.m_total_blocks = n; p.m_perceptual = perceptual;
	for_each_item(n, threads, [&]() { oclref::refine_endpoint_clusterization(p, (const oclref::pixel_block*)pBlocks, (const oclref::rec_block_struct*)pBlock_info, (const oclref::rec_endpoint_cluster_struct*)pClusters, pSorted_block_indices, pOut); });
}

// opencl_find_optimal_selector_clusters_for_each_block (opencl.cpp:1108) -> kernel of the same name (cl:1159)
OCL_API void oclref_find_optimal_selector_clusters_for_each_block(const void* pBlocks, uint32_t n, const void* pBlock_info, const void* pSelectors, const uint32_t* pSelector_cluster_indices, uint32_t* pOut, int perceptual, uint32_t threads)
{
	oclref::fosc_param_struct p; p.m_total_blocks = n; p.m_perceptual = perceptual;
	for_each_item(n, threads, [&]() { oclref::find_optimal_selector_clusters_for_each_block(p, (const oclref::pixel_block*)pBlocks, (const oclref::fosc_block_struct*)pBlock_info, (const oclref::fosc_selector_struct*)pSelectors, pSelector_cluster_indices, pOut); });
}

// opencl_determine_selectors (opencl.cpp:1165) -> kernel determine_selectors (cl:1227)
OCL_API void oclref_determine_selectors(const void* pBlocks, uint32_t n, const void* pColor5_inten, void* pOut, int perceptual, uint32_t threads)
{
	oclref::ds_param_struct p; p.m_total_blocks = n; p.m_perceptual = perceptual;
	memset(pOut, 0, (size_t)n * 8);
	for_each_item(n, threads, [&]() { oclref::determine_selectors(p, (const oclref::pixel_block*)pBlocks, (const oclref::color_rgba*)pColor5_inten, (oclref::etc_block*)pOut); });
}
