// integration/basisu_b200_seam.h -- the B200 extensions of the reference's GPU seam (encoder/basisu_opencl.h).
//
// The stock seam covers five per-block ETC1S stages. These entry points widen it to the rest of the hot path; they take the
// reference's own opencl_context_ptr (owned by basis_compressor, encoder/basisu_comp.cpp:306-309, 671-675, and handed to the
// frontend in basisu_frontend::params), so no class layout changes. They are declared here, defined in
// integration/basisu_opencl_b200.cpp on top of the C ABI (include/basisu_b200.h), and called from the reference sources
// through the patches under integration/patches/ (applied to copies by integration/Makefile). Semantics are the seam's:
// false => the caller sets m_opencl_failed and runs its own CPU code; outputs are untouched on failure.
#pragma once
#include "basisu_opencl.h"
#include "basisu_uastc_enc.h"
#include "basisu_etc.h"

namespace basisu
{
	// encode_slices_to_uastc_4x4_ldr (encoder/basisu_comp.cpp:1973-2064): image::extract_block_clamped + encode_uastc for every
	// 4x4 block of a slice. pImage = width x height texels, pitch_in_pixels apart; pDst_blocks receives ceil(w/4)*ceil(h/4)
	// basist::uastc_block in raster block order (gpu_image's layout).
	bool opencl_b200_encode_uastc_image(opencl_context_ptr pContext, const color_rgba* pImage, uint32_t width, uint32_t height, uint32_t pitch_in_pixels,
		void* pDst_blocks, uint32_t uastc_flags);

	// uastc_rdo (encoder/basisu_uastc_enc.h:139, called at comp.cpp:2076), in place, same total_jobs chain split.
	bool opencl_b200_uastc_rdo(opencl_context_ptr pContext, uint32_t num_blocks, basist::uastc_block* pBlocks, const color_rgba* pBlock_pixels,
		const uastc_rdo_params& params, uint32_t flags, uint32_t total_jobs);

	// generate_hierarchical_codebook_threaded (encoder/basisu_enc.h:2219) for tree_vector_quant<vec6F> (dim 6: endpoints,
	// frontend.cpp:868) and <vec16F> (dim 16: selectors, frontend.cpp:2140). pTraining_vecs points at the quantizer's
	// std::pair<vecNF, uint64_t> array (stride / weight offset given); codebook and parent_codebook are filled exactly as the
	// CPU function fills them (clusters of training-vector indices, same order).
	bool opencl_b200_generate_hierarchical_codebook(opencl_context_ptr pContext, uint32_t dim, const void* pTraining_vecs, uint32_t num_training_vecs,
		size_t stride_bytes, size_t weight_offset_bytes, uint32_t max_codebook_size, uint32_t max_parent_codebook_size,
		basisu::vector<uint_vec>& codebook, basisu::vector<uint_vec>& parent_codebook, uint32_t max_threads, bool even_odd_input_pairs_equal);

	// generate_endpoint_codebook, step 0 (frontend.cpp:1214-1610): clusters as CSR lists of block indices into the array given to
	// opencl_set_pixel_blocks; one etc_block (base colour + intensity table) out per cluster.
	bool opencl_b200_encode_etc1s_endpoint_clusters(opencl_context_ptr pContext, etc_block* pOutput_blocks, uint32_t total_clusters,
		const uint32_t* pCluster_offsets, const uint32_t* pCluster_block_indices, bool perceptual, uint32_t total_perms);

	// create_optimized_selector_codebook (frontend.cpp:2259-2345): per cluster the 16 optimised selectors, texel (x, y) at bits
	// 2 * (x + 4 * y) of pOutput_selectors[cluster]; pEtc_blocks = m_encoded_blocks (all blocks of the slice).
	bool opencl_b200_optimize_selector_codebook(opencl_context_ptr pContext, const etc_block* pEtc_blocks, uint32_t total_clusters,
		const uint32_t* pCluster_offsets, const uint32_t* pCluster_block_indices, uint32_t* pOutput_selectors, bool perceptual);

	// reoptimize_remapped_endpoints' per-cluster loop (frontend.cpp:3008-3090), called by the backend after endpoint remapping:
	// pBlock_selectors[k] = packed selectors of the k-th listed block (texel (x, y) at bits 2 * (x + 4 * y)); endpoints travel as
	// color_rgba(r5, g5, b5, intensity table). The caller applies `new_err < cur_err`.
	bool opencl_b200_reoptimize_endpoint_clusters(opencl_context_ptr pContext, uint32_t total_clusters, const uint32_t* pCluster_offsets,
		const uint32_t* pCluster_block_indices, const uint32_t* pBlock_selectors, const color_rgba* pCluster_color5_inten,
		color_rgba* pNew_color5_inten, uint64_t* pNew_err, uint64_t* pCur_err, bool perceptual, uint32_t total_perms);

	// (pBlock_selectors == nullptr: free selectors = generate_endpoint_codebook at refinement steps >= 1, frontend.cpp:1493-1606,
	// with pCluster_color5_inten the previous endpoints and pCur_err their error.)

	// compute_endpoint_subblock_error_vec (frontend.cpp:1006-1082): pOut_errors[2 * block + subblock] for every source block, against
	// the endpoint of the block's cluster (pBlock_color5_inten: color_rgba(r5, g5, b5, table) per block).
	bool opencl_b200_compute_subblock_errors(opencl_context_ptr pContext, const color_rgba* pBlock_color5_inten, uint64_t* pOut_errors, bool perceptual);

	// basisu_backend::create_encoder_blocks' endpoint prediction / endpoint RDO scan (backend.cpp:405-617, non-video): slices as
	// (first block, blocks per row, rows) triples, the frontend's output blocks and endpoint codebook in, the decided endpoint index and
	// predictor per block out (predictor 3 = none; bit 7 on a 3: the block's error was zero, which the statistics skip).
	bool opencl_b200_backend_endpoint_prediction(opencl_context_ptr pContext, uint32_t num_slices, const uint32_t* pSlice_first_block_nbx_nby, const etc_block* pEtc_blocks,
		uint32_t total_endpoints, const color_rgba* pEndpoint_color5_inten, float endpoint_rdo_quality_thresh, bool perceptual, uint32_t* pBlock_endpoint_indices, uint8_t* pOut_predictors);

	// basisu::image_resample for 8-bit images (enc.cpp:1022-1171), as basis_compressor::generate_mipmaps calls it per mip level
	// (comp.cpp:2203-2218): same arguments, same bytes in dst. The contributor lists come from the reference's own Resampler.
	bool opencl_b200_image_resample(opencl_context_ptr pContext, const image& src, image& dst, bool srgb, const char* pFilter, float filter_scale, bool wrapping,
		uint32_t first_comp, uint32_t num_comps);

	// palette_index_reorderer::init(num_indices, pIndices, num_syms, nullptr, nullptr, 0) + get_remap_table() (enc.cpp:1785-1915) for
	// the endpoint palette (backend.cpp:196-198): remap_table gets the same old -> new table.
	bool opencl_b200_palette_reorder(opencl_context_ptr pContext, uint32_t num_indices, const uint32_t* pIndices, uint32_t num_syms, uint_vec& remap_table);

	// Stage clocks of the patched compressor (basis_compressor::process): name -> seconds of the last run, readable from outside
	// through `extern "C" double b200_dropin_stage_secs(const char* name)` (benchmarks; no effect on the output).
	void opencl_b200_note_stage_secs(const char* pName, double secs);
} // namespace basisu
