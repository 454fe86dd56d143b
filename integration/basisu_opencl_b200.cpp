// integration/basisu_opencl_b200.cpp -- drop-in replacement for the reference's encoder/basisu_opencl.cpp.
//
// Implements the reference's GPU seam (encoder/basisu_opencl.h:24-141, declared there, unchanged) on top of the C ABI of
// libbasisu_b200.so (include/basisu_b200.h).  Build the reference's encoder library with this file instead of
// encoder/basisu_opencl.cpp and link libbasisu_b200.so: `basisu -opencl`, cFlagUseOpenCL and m_use_opencl then run the
// ETC1S frontend's five per-block stages on the B200, with zero changes to basisu_frontend.cpp or basisu_tool.cpp.
// Failure semantics are the reference's: any call returning false makes the frontend null its context, set
// m_opencl_failed and recompute that stage on the CPU (encoder/basisu_frontend.cpp:757-762).
#include "basisu_b200_seam.h"
#include "basisu_resampler.h"
#include "basisu_b200.h"
#include <map>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <string>

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

	void opencl_b200_note_stage_secs(const char* pName, double secs);

	opencl_context_ptr opencl_create_context()
	{
		if (!g_b200_available) return nullptr;
		interval_timer tm; tm.start();
		struct note { interval_timer& t; ~note() { opencl_b200_note_stage_secs("context_create", t.get_elapsed_secs()); } } n{ tm };
		// One process per GPU: B200_DEVICE selects the device (default 0); B200_COMM_WORLD / B200_COMM_RANK / B200_COMM_ID (the 128
		// bytes of b200_comm_unique_id in hex, distributed by the launcher) attach the NCCL communicator that merges stage outputs.
		const char* pDev = getenv("B200_DEVICE");
		b200_context* c = b200_create_context(pDev ? atoi(pDev) : 0);
		if (!c) { error_printf("opencl_create_context (b200): %s\n", b200_last_error(nullptr)); return nullptr; }
		const char* pWorld = getenv("B200_COMM_WORLD"); const char* pRank = getenv("B200_COMM_RANK"); const char* pId = getenv("B200_COMM_ID");
		if (pWorld && atoi(pWorld) > 1)
		{
			uint8_t id[128];
			bool ok = pRank && pId && strlen(pId) == 256;
			for (uint32_t i = 0; ok && i < 128; i++) { unsigned v = 0; ok = sscanf(pId + i * 2, "%2x", &v) == 1; id[i] = (uint8_t)v; }
			if (!ok || !b200_comm_init(c, atoi(pRank), atoi(pWorld), id))
			{
				error_printf("opencl_create_context (b200): communicator setup failed: %s\n", ok ? b200_last_error(c) : "B200_COMM_RANK / B200_COMM_ID missing or malformed");
				b200_destroy_context(c);
				return nullptr;
			}
		}
		opencl_context* p = new opencl_context;
		p->m_ctx = c;
		return p;
	}

	void opencl_destroy_context(opencl_context_ptr context)
	{
		if (!context) return;
		interval_timer tm; tm.start();
		b200_destroy_context(context->m_ctx);
		delete context;
		opencl_b200_note_stage_secs("context_destroy", tm.get_elapsed_secs());
	}

	static bool report(opencl_context_ptr p, int ok, const char* what)
	{
		if (!ok) error_printf("%s (b200): %s\n", what, b200_last_error(p ? p->m_ctx : nullptr));
		else if (p) debug_printf("[b200] %s: %3.3f ms of kernels, %u launches\n", what, b200_last_kernel_ms(p->m_ctx), b200_last_launch_count(p->m_ctx));
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

	// ---- extensions of the seam (integration/basisu_b200_seam.h) ----------------------------------------------------------

	bool opencl_b200_encode_uastc_image(opencl_context_ptr p, const color_rgba* pImage, uint32_t width, uint32_t height, uint32_t pitch_in_pixels,
		void* pDst_blocks, uint32_t uastc_flags)
	{
		if (!p) return false;
		return report(p, b200_uastc_encode_image(p->m_ctx, pImage, width, height, (size_t)pitch_in_pixels * sizeof(color_rgba), pDst_blocks, uastc_flags), "opencl_b200_encode_uastc_image");
	}

	bool opencl_b200_uastc_rdo(opencl_context_ptr p, uint32_t num_blocks, basist::uastc_block* pBlocks, const color_rgba* pBlock_pixels,
		const uastc_rdo_params& params, uint32_t flags, uint32_t total_jobs)
	{
		if (!p) return false;
		b200_uastc_rdo_params q;
		q.lz_dict_size = params.m_lz_dict_size; q.lambda = params.m_lambda;
		q.max_allowed_rms_increase_ratio = params.m_max_allowed_rms_increase_ratio;
		q.skip_block_rms_thresh = params.m_skip_block_rms_thresh;
		q.endpoint_refinement = params.m_endpoint_refinement ? 1u : 0u;
		q.max_smooth_block_std_dev = params.m_max_smooth_block_std_dev;
		q.smooth_block_max_error_scale = params.m_smooth_block_max_error_scale;
		q.lz_literal_cost = params.m_lz_literal_cost;
		return report(p, b200_uastc_rdo(p->m_ctx, num_blocks, pBlocks, pBlock_pixels, &q, flags, total_jobs), "opencl_b200_uastc_rdo");
	}

	bool opencl_b200_generate_hierarchical_codebook(opencl_context_ptr p, uint32_t dim, const void* pTraining_vecs, uint32_t num_training_vecs,
		size_t stride_bytes, size_t weight_offset_bytes, uint32_t max_codebook_size, uint32_t max_parent_codebook_size,
		basisu::vector<uint_vec>& codebook, basisu::vector<uint_vec>& parent_codebook, uint32_t max_threads, bool even_odd_input_pairs_equal)
	{
		if (!p) return false;
		b200_tsvq_result r;
		if (!report(p, b200_tsvq_generate(p->m_ctx, dim, num_training_vecs, pTraining_vecs, stride_bytes, weight_offset_bytes, max_codebook_size, max_parent_codebook_size,
			max_threads, even_odd_input_pairs_equal, &r), "opencl_b200_generate_hierarchical_codebook"))
			return false;
		debug_printf("b200_tsvq_generate: dim %u, %u training vectors, %u unique, %u clusters, %u parent clusters, %u device rounds, %u node splits, %3.3f ms on the device\n",
			dim, num_training_vecs, r.num_unique, r.num_clusters, r.num_parent_clusters, r.rounds, r.nodes_split, b200_last_kernel_ms(p->m_ctx));
		codebook.resize(0);
		codebook.resize(r.num_clusters);
		for (uint32_t i = 0; i < r.num_clusters; i++)
		{
			const uint32_t n = r.cluster_offsets[i + 1] - r.cluster_offsets[i];
			codebook[i].resize(n);
			if (n) memcpy(codebook[i].data(), r.cluster_indices + r.cluster_offsets[i], (size_t)n * sizeof(uint32_t));
		}
		parent_codebook.resize(0);
		parent_codebook.resize(r.num_parent_clusters);
		for (uint32_t i = 0; i < r.num_parent_clusters; i++)
		{
			const uint32_t n = r.parent_offsets[i + 1] - r.parent_offsets[i];
			parent_codebook[i].resize(n);
			if (n) memcpy(parent_codebook[i].data(), r.parent_indices + r.parent_offsets[i], (size_t)n * sizeof(uint32_t));
		}
		return true;
	}

	bool opencl_b200_encode_etc1s_endpoint_clusters(opencl_context_ptr p, etc_block* pOutput_blocks, uint32_t total_clusters,
		const uint32_t* pCluster_offsets, const uint32_t* pCluster_block_indices, bool perceptual, uint32_t total_perms)
	{
		if (!p) return false;
		return report(p, b200_etc1s_encode_endpoint_clusters(p->m_ctx, pOutput_blocks, total_clusters, pCluster_offsets, pCluster_block_indices, perceptual, total_perms),
			"opencl_b200_encode_etc1s_endpoint_clusters");
	}

	bool opencl_b200_optimize_selector_codebook(opencl_context_ptr p, const etc_block* pEtc_blocks, uint32_t total_clusters,
		const uint32_t* pCluster_offsets, const uint32_t* pCluster_block_indices, uint32_t* pOutput_selectors, bool perceptual)
	{
		if (!p) return false;
		return report(p, b200_etc1s_optimize_selector_codebook(p->m_ctx, pEtc_blocks, total_clusters, pCluster_offsets, pCluster_block_indices, pOutput_selectors, perceptual),
			"opencl_b200_optimize_selector_codebook");
	}

	bool opencl_b200_reoptimize_endpoint_clusters(opencl_context_ptr p, uint32_t total_clusters, const uint32_t* pCluster_offsets,
		const uint32_t* pCluster_block_indices, const uint32_t* pBlock_selectors, const color_rgba* pCluster_color5_inten,
		color_rgba* pNew_color5_inten, uint64_t* pNew_err, uint64_t* pCur_err, bool perceptual, uint32_t total_perms)
	{
		if (!p) return false;
		return report(p, b200_etc1s_reoptimize_endpoint_clusters(p->m_ctx, total_clusters, pCluster_offsets, pCluster_block_indices, pBlock_selectors, pCluster_color5_inten,
			pNew_color5_inten, pNew_err, pCur_err, perceptual, total_perms), "opencl_b200_reoptimize_endpoint_clusters");
	}

	bool opencl_b200_compute_subblock_errors(opencl_context_ptr p, const color_rgba* pBlock_color5_inten, uint64_t* pOut_errors, bool perceptual)
	{
		if (!p) return false;
		return report(p, b200_etc1s_subblock_errors(p->m_ctx, pBlock_color5_inten, pOut_errors, perceptual), "opencl_b200_compute_subblock_errors");
	}

	bool opencl_b200_backend_endpoint_prediction(opencl_context_ptr p, uint32_t num_slices, const uint32_t* pSlice_first_block_nbx_nby, const etc_block* pEtc_blocks,
		uint32_t total_endpoints, const color_rgba* pEndpoint_color5_inten, float endpoint_rdo_quality_thresh, bool perceptual, uint32_t* pBlock_endpoint_indices, uint8_t* pOut_predictors)
	{
		if (!p) return false;
		return report(p, b200_etc1s_backend_endpoint_prediction(p->m_ctx, num_slices, pSlice_first_block_nbx_nby, pEtc_blocks, total_endpoints, pEndpoint_color5_inten,
			endpoint_rdo_quality_thresh, perceptual, pBlock_endpoint_indices, pOut_predictors), "opencl_b200_backend_endpoint_prediction");
	}

	bool opencl_b200_image_resample(opencl_context_ptr p, const image& src, image& dst, bool srgb, const char* pFilter, float filter_scale, bool wrapping,
		uint32_t first_comp, uint32_t num_comps)
	{
		if (!p) return false;
		const uint32_t src_w = src.get_width(), src_h = src.get_height(), dst_w = dst.get_width(), dst_h = dst.get_height();
		// image_resample's own argument checks (enc.cpp:1035-1059); anything it would reject or short-cut is left to it
		if (!src_w || !src_h || !dst_w || !dst_h || (num_comps < 1) || (first_comp + num_comps > 4)) return false;
		if ((maximum(src_w, src_h) > BASISU_RESAMPLER_MAX_DIMENSION) || (maximum(dst_w, dst_h) > BASISU_RESAMPLER_MAX_DIMENSION)) return false;
		if ((src_w == dst_w) && (src_h == dst_h) && (filter_scale == 1.0f)) return false;

		// The reference's contributor lists, from its own Resampler (what image_resample constructs for component 0, enc.cpp:1080-1083)
		Resampler resampler(src_w, src_h, dst_w, dst_h, wrapping ? Resampler::BOUNDARY_WRAP : Resampler::BOUNDARY_CLAMP, 0.0f, 1.0f, pFilter, nullptr, nullptr,
			filter_scale, filter_scale, 0, 0);
		if (resampler.status() != Resampler::STATUS_OKAY) return false;
		const Resampler::Contrib_List* pClist[2] = { resampler.get_clist_x(), resampler.get_clist_y() };
		const uint32_t n[2] = { dst_w, dst_h };
		std::vector<uint32_t> offsets[2];
		std::vector<b200_resample_contrib> contribs[2];
		for (int axis = 0; axis < 2; axis++)
		{
			offsets[axis].resize(n[axis] + 1);
			for (uint32_t i = 0; i < n[axis]; i++)
			{
				offsets[axis][i] = (uint32_t)contribs[axis].size();
				for (uint32_t j = 0; j < pClist[axis][i].n; j++)
				{
					b200_resample_contrib c;
					c.weight = pClist[axis][i].p[j].weight; c.pixel = pClist[axis][i].p[j].pixel;
					contribs[axis].push_back(c);
				}
			}
			offsets[axis][n[axis]] = (uint32_t)contribs[axis].size();
		}

		// The two sRGB tables exactly as image_resample fills them (enc.cpp:1061-1075)
		static float s_srgb_to_linear[256];
		static uint8_t s_linear_to_srgb[8192];
		static std::once_flag s_tables_once;
		std::call_once(s_tables_once, [] {
			for (int i = 0; i < 256; ++i) s_srgb_to_linear[i] = srgb_to_linear((float)i * (1.0f / 255.0f));
			for (int i = 0; i < 8192; ++i) s_linear_to_srgb[i] = (uint8_t)clamp<int>((int)(255.0f * linear_to_srgb((float)i * (1.0f / (8192 - 1))) + .5f), 0, 255);
		});

		return report(p, b200_image_resample_rgba8(p->m_ctx, src.get_ptr(), src_w, src_h, (size_t)src.get_pitch() * sizeof(color_rgba), dst.get_ptr(), dst_w, dst_h,
			(size_t)dst.get_pitch() * sizeof(color_rgba), offsets[0].data(), contribs[0].data(), offsets[1].data(), contribs[1].data(), first_comp, num_comps,
			srgb ? s_srgb_to_linear : nullptr, srgb ? s_linear_to_srgb : nullptr), "opencl_b200_image_resample");
	}

	bool opencl_b200_palette_reorder(opencl_context_ptr p, uint32_t num_indices, const uint32_t* pIndices, uint32_t num_syms, uint_vec& remap_table)
	{
		if (!p) return false;
		uint_vec table(num_syms);
		if (!report(p, b200_palette_reorder(p->m_ctx, num_indices, pIndices, num_syms, table.data()), "opencl_b200_palette_reorder")) return false;
		remap_table.swap(table);
		return true;
	}

	static std::mutex g_stage_mutex;
	static std::map<std::string, double> g_stage_secs;

	void opencl_b200_note_stage_secs(const char* pName, double secs)
	{
		debug_printf("[stage] %s: %3.3f secs\n", pName, secs);
		std::lock_guard<std::mutex> lk(g_stage_mutex);
		g_stage_secs[pName] = secs;
	}
} // namespace basisu

extern "C" __attribute__((visibility("default"))) double b200_dropin_stage_secs(const char* pName)
{
	std::lock_guard<std::mutex> lk(basisu::g_stage_mutex);
	auto it = basisu::g_stage_secs.find(pName);
	return (it == basisu::g_stage_secs.end()) ? -1.0 : it->second;
}
