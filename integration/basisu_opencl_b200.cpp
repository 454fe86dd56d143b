// integration/basisu_opencl_b200.cpp -- drop-in replacement for the reference's encoder/basisu_opencl.cpp.
//
// Implements the reference's GPU seam (encoder/basisu_opencl.h:24-141, declared there, unchanged) on top of the C ABI of
// libbasisu_b200.so (include/basisu_b200.h).  Build the reference's encoder library with this file instead of
// encoder/basisu_opencl.cpp and link libbasisu_b200.so: `basisu -opencl`, cFlagUseOpenCL and m_use_opencl then run the
// ETC1S frontend's five per-block stages on the B200, with zero changes to basisu_frontend.cpp or basisu_tool.cpp.
// Failure semantics are the reference's: any call returning false makes the frontend null its context, set
// m_opencl_failed and recompute that stage on the CPU (encoder/basisu_frontend.cpp:757-762).
#include "encoder/basisu_opencl.h"
#include "basisu_b200.h"

namespace basisu
{
	static bool g_b200_available = false;

	// opencl_context is an opaque struct in the header; here it simply wraps the b200 context.
	struct opencl_context { b200_context* m_ctx; };

	bool opencl_init(bool force_serialization)
	{
		(void)force_serialization; // serialisation existed for buggy OpenCL drivers (opencl.cpp:690-708); CUDA contexts are independent
		g_b200_available = b200_device_count() > 0;
		return g_b200_available;
	}

	void opencl_deinit() { g_b200_available = false; }
	bool opencl_is_available() { return g_b200_available; }

	opencl_context_ptr opencl_create_context()
	{
		if (!g_b200_available) return nullptr;
		b200_context* c = b200_create_context(0);
		if (!c) { error_printf("opencl_create_context (b200): %s\n", b200_last_error(nullptr)); return nullptr; }
		opencl_context* p = new opencl_context;
		p->m_ctx = c;
		return p;
	}

	void opencl_destroy_context(opencl_context_ptr context)
	{
		if (!context) return;
		b200_destroy_context(context->m_ctx);
		delete context;
	}

	static bool report(opencl_context_ptr p, int ok, const char* what)
	{
		if (!ok) error_printf("%s (b200): %s\n", what, b200_last_error(p ? p->m_ctx : nullptr));
		return ok != 0;
	}

	bool opencl_set_pixel_blocks(opencl_context_ptr p, size_t total_blocks, const cl_pixel_block* pPixel_blocks)
	{
		if (!p) return false;
		return report(p, b200_etc1s_set_pixel_blocks(p->m_ctx, (uint32_t)total_blocks, pPixel_blocks), "opencl_set_pixel_blocks");
	}

	bool opencl_encode_etc1s_blocks(opencl_context_ptr p, etc_block* pOutput_blocks, bool perceptual, uint32_t total_perms)
	{
		if (!p) return false;
		return report(p, b200_etc1s_encode_blocks(p->m_ctx, pOutput_blocks, perceptual, total_perms), "opencl_encode_etc1s_blocks");
	}

	bool opencl_encode_etc1s_pixel_clusters(opencl_context_ptr p, etc_block* pOutput_blocks, uint32_t total_clusters, const cl_pixel_cluster* pClusters,
		uint64_t total_pixels, const color_rgba* pPixels, const uint32_t* pPixel_weights, bool perceptual, uint32_t total_perms)
	{
		if (!p) return false;
		static_assert(sizeof(cl_pixel_cluster) == sizeof(b200_pixel_cluster), "layout");
		return report(p, b200_etc1s_encode_pixel_clusters(p->m_ctx, pOutput_blocks, total_clusters, reinterpret_cast<const b200_pixel_cluster*>(pClusters),
			total_pixels, pPixels, pPixel_weights, perceptual, total_perms), "opencl_encode_etc1s_pixel_clusters");
	}

	bool opencl_refine_endpoint_clusterization(opencl_context_ptr p, const cl_block_info_struct* pPixel_block_info, uint32_t total_clusters,
		const cl_endpoint_cluster_struct* pCluster_info, const uint32_t* pSorted_block_indices, uint32_t* pOutput_cluster_indices, bool perceptual)
	{
		if (!p) return false;
		static_assert(sizeof(cl_block_info_struct) == sizeof(b200_block_info) && sizeof(cl_endpoint_cluster_struct) == sizeof(b200_endpoint_cluster), "layout");
		return report(p, b200_etc1s_refine_endpoint_clusterization(p->m_ctx, reinterpret_cast<const b200_block_info*>(pPixel_block_info), total_clusters,
			reinterpret_cast<const b200_endpoint_cluster*>(pCluster_info), pSorted_block_indices, pOutput_cluster_indices, perceptual), "opencl_refine_endpoint_clusterization");
	}

	bool opencl_find_optimal_selector_clusters_for_each_block(opencl_context_ptr p, const fosc_block_struct* pInput_block_info, uint32_t total_input_selectors,
		const fosc_selector_struct* pInput_selectors, const uint32_t* pSelector_cluster_indices, uint32_t* pOutput_selector_cluster_indices, bool perceptual)
	{
		if (!p) return false;
		static_assert(sizeof(fosc_block_struct) == sizeof(b200_fosc_block) && sizeof(fosc_selector_struct) == sizeof(b200_fosc_selector), "layout");
		return report(p, b200_etc1s_find_optimal_selector_clusters_for_each_block(p->m_ctx, reinterpret_cast<const b200_fosc_block*>(pInput_block_info), total_input_selectors,
			reinterpret_cast<const b200_fosc_selector*>(pInput_selectors), pSelector_cluster_indices, pOutput_selector_cluster_indices, perceptual),
			"opencl_find_optimal_selector_clusters_for_each_block");
	}

	bool opencl_determine_selectors(opencl_context_ptr p, const color_rgba* pInput_etc_color5_and_inten, etc_block* pOutput_blocks, bool perceptual)
	{
		if (!p) return false;
		return report(p, b200_etc1s_determine_selectors(p->m_ctx, pInput_etc_color5_and_inten, pOutput_blocks, perceptual), "opencl_determine_selectors");
	}
} // namespace basisu
