/* basisu_b200.h -- C ABI of libbasisu_b200.so: the B200 (sm_100a) implementation of basis_universal's per-block
 * encode hot path (UASTC LDR 4x4 and the ETC1S frontend's per-block stages).
 *
 * Plain pointers and sizes only; no C++ or torch types cross this boundary.  Every entry point names the reference
 * interface it replaces (paths relative to the BinomialLLC/basis_universal tree); INTEGRATION.md shows the call-site
 * patches.  Block layouts are the reference's own:
 *   source block  = basisu::pixel_block, 64 B, [y][x] RGBA8, R first          (encoder/basisu_enc.h:4156)
 *   UASTC block   = basist::uastc_block, 16 B                                   (transcoder/basisu_transcoder_uastc.h:198)
 *   ETC1 block    = basisu::etc_block, 8 B, big-endian bitfield                 (encoder/basisu_etc.h:91)
 *
 * Error model mirrors the reference's OpenCL seam (encoder/basisu_opencl.h:24-141): functions return 1 on success and
 * 0 on failure; on failure the caller may fall back to its own CPU code.  b200_last_error() describes the failure.
 * There is NO CPU fallback inside this library: without a usable CUDA device every call fails loudly.
 */
#ifndef BASISU_B200_H
#define BASISU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_context b200_context;

/* Number of usable CUDA devices (>= 0), or -1 if the CUDA runtime cannot be initialised.
 * Replaces: basisu::opencl_init / opencl_is_available (encoder/basisu_opencl.h:24-27). */
int b200_device_count(void);

/* One context = one device + one stream + its scratch buffers; use one per host thread.
 * Replaces: opencl_create_context / opencl_destroy_context (encoder/basisu_opencl.h:30-33). */
b200_context* b200_create_context(int device_index);
void b200_destroy_context(b200_context* ctx);

/* Human-readable description of the last failure on this context (never NULL). ctx may be NULL for creation failures. */
const char* b200_last_error(const b200_context* ctx);

/* ---- UASTC LDR 4x4 ------------------------------------------------------------------------------------------------ */

/* Encodes num_blocks source blocks to UASTC. Batch form of
 *     void basisu::encode_uastc(const uint8_t* pRGBAPixels, basist::uastc_block& out, uint32_t flags)
 * (encoder/basisu_uastc_enc.h:68), replacing the job-pool loop at encoder/basisu_comp.cpp:2006-2064.
 * `flags` is the reference's flag word unchanged: low 3 bits = cPackUASTCLevel*, plus the cPackUASTCFavor... and cPackUASTCETC1... bits
 * (encoder/basisu_uastc_enc.h:24-63).  Output bytes are bit-identical to the reference's.
 * pBlocks / pOut are HOST pointers; copies to and from the device happen inside the call. */
int b200_uastc_encode_blocks(b200_context* ctx, const void* pBlocks, uint32_t num_blocks, void* pOut, uint32_t flags);

/* Same, with DEVICE pointers (inputs already resident in HBM, outputs left there). Work is enqueued on the context's
 * stream and the call returns after the stream has been synchronised. */
int b200_uastc_encode_blocks_device(b200_context* ctx, const void* dBlocks, uint32_t num_blocks, void* dOut, uint32_t flags);

/* Raster in, blocks out: basis_compressor::extract_source_blocks (encoder/basisu_comp.cpp:3207) followed by the encode loop
 * (comp.cpp:2006-2064) for one slice, without the intermediate pixel_block array ever existing on the host. pRGBA is a HOST
 * pointer to height rows of width RGBA8 texels, pitch_bytes apart; pOut receives ((width+3)/4) * ((height+3)/4) blocks. */
int b200_uastc_encode_image(b200_context* ctx, const void* pRGBA, uint32_t width, uint32_t height, size_t pitch_bytes, void* pOut, uint32_t flags);

/* RDO post-pass, in place. Same contract as
 *     bool basisu::uastc_rdo(uint32_t num_blocks, basist::uastc_block* pBlocks, const color_rgba* pBlock_pixels,
 *                            const uastc_rdo_params& params, uint32_t flags, job_pool*, uint32_t total_jobs)
 * (encoder/basisu_uastc_enc.h:139; called at encoder/basisu_comp.cpp:2076). `total_jobs` chunks the block range
 * exactly as the reference does (uastc_enc.cpp:4116-4154), because the result depends on it. */
typedef struct b200_uastc_rdo_params
{
	uint32_t lz_dict_size;                /* uastc_rdo_params::m_lz_dict_size */
	float lambda;                         /* m_lambda */
	float max_allowed_rms_increase_ratio; /* m_max_allowed_rms_increase_ratio */
	float skip_block_rms_thresh;          /* m_skip_block_rms_thresh */
	uint32_t endpoint_refinement;         /* m_endpoint_refinement */
	float max_smooth_block_std_dev;       /* m_max_smooth_block_std_dev */
	float smooth_block_max_error_scale;   /* m_smooth_block_max_error_scale */
	uint32_t lz_literal_cost;             /* m_lz_literal_cost */
} b200_uastc_rdo_params;
int b200_uastc_rdo(b200_context* ctx, uint32_t num_blocks, void* pBlocks, const void* pBlock_pixels,
	const b200_uastc_rdo_params* params, uint32_t flags, uint32_t total_jobs);
/* On failure (0) pBlocks is left untouched, so the caller may run its own CPU uastc_rdo on the encoder's output. */

/* Many slices in one call: pSlice_num_blocks[s] blocks per slice, the slices' blocks (and source texels) laid end to end.
 * Equivalent to calling uastc_rdo once per slice with the same parameters and total_jobs (the loop at comp.cpp:1996-2089 over
 * slices), but every chain of every slice runs concurrently (a chain is sequential; chains and slices are independent).
 * The _device form takes DEVICE pointers (e.g. the output of b200_uastc_encode_blocks_device) and leaves the result there. */
int b200_uastc_rdo_batch(b200_context* ctx, uint32_t num_slices, const uint32_t* pSlice_num_blocks, void* pBlocks, const void* pBlock_pixels,
	const b200_uastc_rdo_params* params, uint32_t flags, uint32_t total_jobs);
int b200_uastc_rdo_batch_device(b200_context* ctx, uint32_t num_slices, const uint32_t* pSlice_num_blocks, void* dBlocks, const void* dBlock_pixels,
	const b200_uastc_rdo_params* params, uint32_t flags, uint32_t total_jobs);

/* encode_slices_to_uastc_4x4_ldr's per-slice body (comp.cpp:1996-2089) for a list of slices: encode_uastc over every block, then,
 * if params is not NULL, uastc_rdo per slice with the same total_jobs. HOST pointers: source blocks in (64 B each, the slices end
 * to end), final UASTC blocks out; the encoded blocks stay in HBM between the two stages. */
int b200_uastc_encode_rdo_blocks(b200_context* ctx, uint32_t num_slices, const uint32_t* pSlice_num_blocks, const void* pBlocks, void* pOut,
	uint32_t flags, const b200_uastc_rdo_params* params, uint32_t total_jobs);

/* ---- ETC1S frontend per-block stages: the reference's existing GPU seam, symbol for symbol ------------------------------ */
/* Packed argument structs are the reference's (encoder/basisu_opencl.h:36-135), restated in C. */

#pragma pack(push, 1)
typedef struct b200_pixel_cluster { uint64_t total_pixels; uint64_t first_pixel_index; } b200_pixel_cluster; /* cl_pixel_cluster, opencl.h:49 */
typedef struct b200_block_info { uint16_t first_cluster_ofs; uint16_t num_clusters; uint16_t cur_cluster_index; uint8_t cur_cluster_etc_inten; } b200_block_info; /* cl_block_info_struct, opencl.h:70 */
typedef struct b200_endpoint_cluster { uint8_t unscaled_r, unscaled_g, unscaled_b, unscaled_a; uint8_t etc_inten; uint16_t cluster_index; } b200_endpoint_cluster; /* cl_endpoint_cluster_struct, opencl.h:81 */
typedef struct b200_fosc_block { uint8_t etc_r, etc_g, etc_b, etc_a; uint32_t first_selector; uint32_t num_selectors; } b200_fosc_block; /* fosc_block_struct, opencl.h:104 */
typedef struct b200_fosc_selector { uint32_t packed_selectors; } b200_fosc_selector; /* fosc_selector_struct, opencl.h:98 */
#pragma pack(pop)

/* The reference has two implementations of the block/cluster optimiser behind this seam: its OpenCL kernels
 * (bin/ocl_kernels.cl:772-982, always prune intensity tables) and the CPU etc1_optimizer the frontend falls back to
 * (encoder/basisu_etc.cpp:948-995, 1104-1278). b200_etc1s_encode_blocks / _pixel_clusters can reproduce either, bit for bit:
 *   CPU_OPTIMIZER (default): output identical to the reference CPU encoder for these stages (total_perms 16/64/165), hence
 *                            identical ETC1S PSNR at the same -q; total_perms 4 (CPU quality "Fast") uses the kernel flavour.
 *   OPENCL_KERNELS:          output identical to the reference's OpenCL path.
 * The three integer stages (refine, selector search, determine_selectors) are the same in both. */
enum { B200_ETC1S_FLAVOUR_OPENCL_KERNELS = 0, B200_ETC1S_FLAVOUR_CPU_OPTIMIZER = 1 };
int b200_etc1s_set_flavour(b200_context* ctx, int flavour);

/* opencl_set_pixel_blocks (opencl.h:46): uploads the slice's source blocks once; later calls read them from HBM. */
int b200_etc1s_set_pixel_blocks(b200_context* ctx, uint32_t total_blocks, const void* pPixel_blocks);
/* opencl_encode_etc1s_blocks (opencl.h:47): per-block ETC1S optimisation, total_perms in {4,16,64,165}. */
int b200_etc1s_encode_blocks(b200_context* ctx, void* pOutput_blocks, int perceptual, uint32_t total_perms);
/* opencl_encode_etc1s_pixel_clusters (opencl.h:58). */
int b200_etc1s_encode_pixel_clusters(b200_context* ctx, void* pOutput_blocks, uint32_t total_clusters, const b200_pixel_cluster* pClusters,
	uint64_t total_pixels, const void* pPixels, const uint32_t* pPixel_weights, int perceptual, uint32_t total_perms);
/* opencl_refine_endpoint_clusterization (opencl.h:90). */
int b200_etc1s_refine_endpoint_clusterization(b200_context* ctx, const b200_block_info* pPixel_block_info, uint32_t total_clusters,
	const b200_endpoint_cluster* pCluster_info, const uint32_t* pSorted_block_indices, uint32_t* pOutput_cluster_indices, int perceptual);
/* opencl_find_optimal_selector_clusters_for_each_block (opencl.h:113). */
int b200_etc1s_find_optimal_selector_clusters_for_each_block(b200_context* ctx, const b200_fosc_block* pInput_block_info, uint32_t total_input_selectors,
	const b200_fosc_selector* pInput_selectors, const uint32_t* pSelector_cluster_indices, uint32_t* pOutput_selector_cluster_indices, int perceptual);
/* opencl_determine_selectors (opencl.h:137). */
int b200_etc1s_determine_selectors(b200_context* ctx, const void* pInput_etc_color5_and_inten, void* pOutput_blocks, int perceptual);

/* ---- ETC1S codebook stages: the wider seam (no OpenCL counterpart in the reference) ------------------------------------------ */

/* Tree-structured vector quantisation of a weighted training set: the device form of
 *     template<typename Quantizer> bool basisu::generate_hierarchical_codebook_threaded(Quantizer& q, uint32_t max_codebook_size,
 *         uint32_t max_parent_codebook_size, basisu::vector<uint_vec>& codebook, basisu::vector<uint_vec>& parent_codebook,
 *         uint32_t max_threads, job_pool*, bool even_odd_input_pairs_equal)                      (encoder/basisu_enc.h:2219)
 * with Quantizer = tree_vector_quant<vec6F> (endpoints, frontend.cpp:868) or <vec16F> (selectors, frontend.cpp:2140), i.e.
 * dim = 6 or 16. Same algorithm: identical training vectors are merged in lexicographic order with summed weights
 * (enc.h:2228-2260); the tree is grown best-first by node variance (tree_vector_quant::generate, enc.h:1631-1676), each split
 * along the principal axis (enc.h:1802-1846) and refined by <= 6 k-means passes (enc.h:1962-2095); with >= 2^18 unique vectors
 * and max_threads > 1 the two-level scheme of enc.h:2112-2214 (max_threads top clusters, one sub-tree each) is followed.
 * The sums over a node's members are parallel reductions in a fixed order (deterministic, but not the reference's serial float
 * order), so clusterings agree with the CPU up to float rounding; the ETC1S gate is +-0.02 dB PSNR, not bit-exactness.
 *
 * pTraining: HOST pointer to num_training records `stride_bytes` apart, each holding `dim` floats at offset 0 and a uint64_t
 * weight at offset `weight_offset_bytes` (the layout of std::pair<vecNF, uint64_t>, passed as is).
 * The result arrays are owned by the context and stay valid until the next b200_tsvq_generate call on it: clusters in the
 * reference's order (leaf nodes by node index), each listing training-vector indices in the reference's order. */
typedef struct b200_tsvq_result
{
	uint32_t num_unique;             /* unique training vectors after merging duplicates */
	uint32_t num_clusters;           /* codebook.size() */
	const uint32_t* cluster_offsets; /* num_clusters + 1 */
	const uint32_t* cluster_indices; /* num_training training-vector indices */
	uint32_t num_parent_clusters;    /* parent_codebook.size(); 0 if max_parent_codebook_size == 0 */
	const uint32_t* parent_offsets;  /* num_parent_clusters + 1 */
	const uint32_t* parent_indices;
	uint32_t rounds, nodes_split;    /* instrumentation: device rounds and node splits computed (speculative ones included) */
} b200_tsvq_result;
int b200_tsvq_generate(b200_context* ctx, uint32_t dim, uint32_t num_training, const void* pTraining, size_t stride_bytes, size_t weight_offset_bytes,
	uint32_t max_codebook_size, uint32_t max_parent_codebook_size, uint32_t max_threads, int even_odd_input_pairs_equal, b200_tsvq_result* pResult);

/* basisu_frontend::generate_endpoint_codebook, step 0 (encoder/basisu_frontend.cpp:1214-1610): one ETC1S colour + intensity
 * table per endpoint cluster, optimised over every texel of the cluster's blocks. Clusters arrive as CSR lists of block indices
 * into the array given to b200_etc1s_set_pixel_blocks, so the host neither gathers nor de-duplicates texels (the reference's
 * OpenCL path radix-sorts every cluster's texels on one host thread first, frontend.cpp:1250-1445). Output: one etc_block per
 * cluster carrying base colour and intensity table, as opencl_encode_etc1s_pixel_clusters writes them. Same flavours as
 * b200_etc1s_encode_pixel_clusters; the CPU_OPTIMIZER result equals the CPU etc1_optimizer's on the gathered texels. */
int b200_etc1s_encode_endpoint_clusters(b200_context* ctx, void* pOutput_blocks, uint32_t total_clusters, const uint32_t* pCluster_offsets,
	const uint32_t* pCluster_block_indices, int perceptual, uint32_t total_perms);

/* basisu_frontend::create_optimized_selector_codebook (encoder/basisu_frontend.cpp:2259-2345): for every selector cluster and
 * each of the 16 texels, the selector (0..3) minimising the summed colour error over the cluster's blocks, each block decoded
 * with its own endpoint (pEtc_blocks: the frontend's m_encoded_blocks, 8 B each, all blocks of the slice). Output: one u32 per
 * cluster, texel (x, y) at bits 2 * (x + 4 * y); 0 for an empty cluster (the reference leaves those entries untouched). */
int b200_etc1s_optimize_selector_codebook(b200_context* ctx, const void* pEtc_blocks, uint32_t total_clusters, const uint32_t* pCluster_offsets,
	const uint32_t* pCluster_block_indices, uint32_t* pOutput_selectors, int perceptual);

/* basisu_frontend::reoptimize_remapped_endpoints, the per-cluster optimiser loop (encoder/basisu_frontend.cpp:3008-3090), which
 * the ETC1S backend calls after its endpoint remapping (basisu_backend.cpp:162, 1284): for every endpoint cluster (CSR lists of
 * block indices, as above) re-fit colour + intensity table over the cluster's texels with each texel's selector IMPOSED
 * (etc1_optimizer::params::m_pForce_selectors; pBlock_selectors[k] = the 16 selectors of the k-th LISTED block, texel (x, y) at
 * bits 2 * (x + 4 * y), i.e. parallel to pCluster_block_indices), and evaluate the cluster's current endpoint
 * (pCluster_color5_inten: r5, g5, b5, table per cluster, 4 B) with the same selectors. total_perms is 64 (cETCQualitySlow) or
 * 165 (cETCQualityUber, compression level 6). Outputs per cluster: the new endpoint (same 4-byte layout), its error and the
 * current error; the caller keeps the reference's `new < current` rule. Empty clusters give zeros.
 * pBlock_selectors == NULL: selectors are free (every texel takes the best of the four colours), which is the per-cluster body of
 * generate_endpoint_codebook at refinement steps >= 1 (encoder/basisu_frontend.cpp:1493-1606): new endpoint + its error + the
 * error of the cluster's previous endpoint; total_perms 16 / 64 / 165 as for b200_etc1s_encode_endpoint_clusters. */
int b200_etc1s_reoptimize_endpoint_clusters(b200_context* ctx, uint32_t total_clusters, const uint32_t* pCluster_offsets, const uint32_t* pCluster_block_indices,
	const uint32_t* pBlock_selectors, const void* pCluster_color5_inten, void* pOut_color5_inten, uint64_t* pOut_new_err, uint64_t* pOut_cur_err,
	int perceptual, uint32_t total_perms);

/* basisu_frontend::compute_endpoint_subblock_error_vec (encoder/basisu_frontend.cpp:1006-1082), the input of
 * introduce_new_endpoint_clusters: for every block of the array given to b200_etc1s_set_pixel_blocks, the error of its two
 * subblocks (flipped layout: texels 0-7 and 8-15) against the endpoint of the block's cluster (pBlock_color5_inten: r5, g5, b5,
 * table per BLOCK, 4 B), best of the four colours per texel. pOut_errors[2 * block + subblock]. The reference evaluates this
 * against the UNSCALED 5-bit base colour plus the table's modifiers (it passes scaled = true, frontend.cpp:1043); so does this. */
int b200_etc1s_subblock_errors(b200_context* ctx, const void* pBlock_color5_inten, uint64_t* pOut_errors, int perceptual);

/* basisu_backend::create_encoder_blocks, the endpoint-prediction / endpoint-RDO pass (encoder/basisu_backend.cpp:405-617; non-video
 * textures): in raster order per slice, a block whose endpoint index equals its left / upper / upper-left neighbour's gets that
 * predictor (0 / 1 / 2); otherwise, if endpoint_rdo_quality_thresh > 0, its index is replaced by the neighbour's whose endpoint
 * keeps the block's error (the block's selectors kept) within max(1, thresh) times its current error, lowest error first,
 * lowest predictor on ties. Neighbour indices are the already-decided ones, exactly as in the reference's scan.
 *   pSlice_first_block_nbx_nby: 3 u32 per slice (first block index into the array given to b200_etc1s_set_pixel_blocks, blocks
 *     per row, rows);  pEtc_blocks: basisu_frontend::get_output_block for every block (8 B, ETC1S: differential mode, delta 0,
 *     both tables equal);  pEndpoint_color5_inten: r5, g5, b5, table per endpoint cluster (4 B);
 *   pBlock_endpoint_indices: in = the frontend's index per block, out = encoder_block::m_endpoint_index;
 *   pOut_predictors: encoder_block::m_endpoint_predictor (3 = NO_ENDPOINT_PRED_INDEX); bit 7 set on a 3 means the block's current
 *     error was zero (the reference's hit / miss statistics leave those out). */
int b200_etc1s_backend_endpoint_prediction(b200_context* ctx, uint32_t num_slices, const uint32_t* pSlice_first_block_nbx_nby, const void* pEtc_blocks,
	uint32_t total_endpoints, const void* pEndpoint_color5_inten, float endpoint_rdo_quality_thresh, int perceptual, uint32_t* pBlock_endpoint_indices, uint8_t* pOut_predictors);

/* basisu::palette_index_reorderer::init(num_indices, pIndices, num_syms, nullptr, nullptr, 0) + get_remap_table()
 * (encoder/basisu_enc.cpp:1785-1915), as basisu_backend::reoptimize_and_sort_endpoints_codebook orders the endpoint palette
 * (encoder/basisu_backend.cpp:196-198): Zeng's greedy ordering over the adjacency counts of the index stream. Same ordering as the
 * reference (same arg-max and tie rules, same float side decision), from sparse adjacency lists instead of the dense
 * num_syms x num_syms table. pRemap_table receives num_syms entries (old index -> new index). HOST pointers. */
int b200_palette_reorder(b200_context* ctx, uint32_t num_indices, const uint32_t* pIndices, uint32_t num_syms, uint32_t* pRemap_table);

/* ---- ETC1S multi-GPU exchange point -------------------------------------------------------------------------------------- */

/* Histogram of the 18-bit endpoint training keys (r5<<13 | g5<<8 | b5<<3 | inten) of `num_blocks` ETC1S blocks, each block
 * counted twice (its two subblocks): a complete description of the shard's endpoint training set
 * (basisu_frontend::init_endpoint_training_vectors, encoder/basisu_frontend.cpp:825-867, followed by the duplicate-merging
 * std::map at the top of generate_hierarchical_codebook_threaded, encoder/basisu_enc.h:2228-2260). Ranks that each own a
 * block-row range SUM all-reduce these 2^18 u32 counters (NCCL) to obtain the global weighted unique training set.
 * Host form: pHist receives 2^18 counters. Device form ACCUMULATES into dHist (caller zeroes it), so the buffer can be
 * handed straight to an all-reduce. */
int b200_etc1s_endpoint_histogram(b200_context* ctx, const void* pEtc_blocks, uint32_t num_blocks, uint32_t* pHist);
int b200_etc1s_endpoint_histogram_device(b200_context* ctx, const void* dEtc_blocks, uint32_t num_blocks, uint32_t* dHist);

/* Per block, the selector training vector of basisu_frontend::generate_selector_clusters (encoder/basisu_frontend.cpp:2156-2179):
 * key = the 16 selectors, 2 bits each, texel (x, y) at bits 2 * (x + 4 * y); weight = clamp(color_distance(low, high block
 * colour) / 300, 1, 4096). The clusterer merges equal vectors, so ranks all-gather their unique (key, summed weight) pairs
 * (basis_universal_b200/distributed.py::allgather_selector_training). */
int b200_etc1s_selector_training(b200_context* ctx, const void* pEtc_blocks, uint32_t num_blocks, int perceptual, uint32_t* pKeys, uint32_t* pWeights);
int b200_etc1s_selector_training_device(b200_context* ctx, const void* dEtc_blocks, uint32_t num_blocks, int perceptual, uint32_t* dKeys, uint32_t* dWeights);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink / NVSwitch ---------------------------------------------------------------- */

/* With a communicator attached, every b200_etc1s_* stage call computes only this rank's share (per-block stages: a contiguous
 * block range = block rows; per-cluster stages: every world-th cluster) and the ranks merge the stage's output array with one
 * NCCL all-reduce, so each rank returns the COMPLETE output: the host logic above (the reference's frontend) runs replicated
 * on identical data. b200_etc1s_set_pixel_blocks still takes the whole slice (268 MB at 8192^2: replicating the texels is what
 * NVSwitch makes cheap; the work is what gets divided). b200_tsvq_generate runs replicated (deterministic, identical on every
 * rank). UASTC needs no collective: shard the blocks by rows and call the encoder on each shard.
 * Rank 0 obtains an id with b200_comm_unique_id and hands its 128 bytes to the other ranks by any means (torch.distributed
 * broadcast, MPI, a file); every rank then calls b200_comm_init on its context. NCCL is loaded with dlopen on first use. */
int b200_comm_unique_id(uint8_t* pId128);
int b200_comm_init(b200_context* ctx, int rank, int world, const uint8_t* pId128);
/* The share of `rank` of n per-block units (host arithmetic, usable without a GPU): [*pFirst, *pLast), contiguous, ceil(n / world) each. */
void b200_shard_range(uint32_t n, uint32_t rank, uint32_t world, uint32_t* pFirst, uint32_t* pLast);
int b200_comm_rank(const b200_context* ctx);
int b200_comm_world(const b200_context* ctx);
/* In-place SUM all-reduce of `count` u32 in device memory over the context's communicator (e.g. the 2^18 endpoint histogram). */
int b200_comm_allreduce_u32_device(b200_context* ctx, void* dBuf, size_t count);
/* Device time, bytes and number of the stage-output merges issued through this context so far. */
int b200_comm_stats(const b200_context* ctx, float* pMs, uint64_t* pBytes, uint32_t* pCalls);
const char* b200_comm_last_error(void);

/* ---- either side of the per-block path: ingest, decode, quality metric ---------------------------------------------------- */

/* Raster RGBA8 image -> array of 64 B pixel_blocks in raster block order, edge texels clamped: one slice of
 * basis_compressor::extract_source_blocks (encoder/basisu_comp.cpp:3207; image::extract_block_clamped,
 * encoder/basisu_enc.h:3168). pitch_bytes is the distance between rows. Writes ((width+3)/4) * ((height+3)/4) blocks. */
int b200_extract_source_blocks(b200_context* ctx, const void* pRGBA, uint32_t width, uint32_t height, size_t pitch_bytes, void* pBlocks);
int b200_extract_source_blocks_device(b200_context* ctx, const void* dRGBA, uint32_t width, uint32_t height, size_t pitch_bytes, void* dBlocks);

/* Decode UASTC LDR 4x4 blocks to 16 RGBA8 texels each (64 B, [y][x] order):
 * bool basist::unpack_uastc(const uastc_block&, color32* pPixels, bool srgb = false) (transcoder/basisu_transcoder.cpp:15886).
 * Returns 0 if any block is invalid (the reference returns false for that block). */
int b200_uastc_unpack_blocks(b200_context* ctx, const void* pUastc_blocks, uint32_t num_blocks, void* pRGBA_blocks);
int b200_uastc_unpack_blocks_device(b200_context* ctx, const void* dUastc_blocks, uint32_t num_blocks, void* dRGBA_blocks);

/* Decode ETC1 blocks (the whole format, of which the frontend's ETC1S blocks are a subset) to 16 RGBA8 texels each:
 * bool basisu::unpack_etc1(const etc_block&, color_rgba* pDst, bool preserve_alpha = false) (encoder/basisu_etc.cpp:604).
 * Returns 0 if a differential base colour overflowed in some block: the reference returns false for such a block without
 * writing its texels; here they are written as zeros, and every valid block is decoded regardless. */
int b200_etc1_unpack_blocks(b200_context* ctx, const void* pEtc1_blocks, uint32_t num_blocks, void* pRGBA_blocks);
int b200_etc1_unpack_blocks_device(b200_context* ctx, const void* dEtc1_blocks, uint32_t num_blocks, void* dRGBA_blocks);

/* The integer part of image_metrics::calc(const image& a, const image& b, ...) (encoder/basisu_enc.cpp:2155) for two images
 * held as block arrays of a width x height image: histograms of |a - b| per channel and of the 709 / 601 luma difference over
 * the texels inside the image (padding texels of edge blocks are not counted), plus the channel sums. The caller finishes
 * with the reference's double arithmetic (basis_universal_b200/image.py::metrics_from_histograms restates it). */
typedef struct b200_block_metrics
{
	uint64_t hist[6][256]; /* R, G, B, A, 709 luma, 601 luma */
	uint64_t sum_a[4], sum_b[4];
} b200_block_metrics;
int b200_block_metrics_device(b200_context* ctx, const void* dBlocksA, const void* dBlocksB, uint32_t width, uint32_t height, b200_block_metrics* pOut);

/* ---- instrumentation ------------------------------------------------------------------------------------------------- */

/* Device time in milliseconds of the kernels of the last successful encode call on this context (CUDA events on the
 * context's stream), and the number of kernel launches it issued. */
float b200_last_kernel_ms(const b200_context* ctx);
uint32_t b200_last_launch_count(const b200_context* ctx);

/* Accumulated per entry-point family since the context was created or b200_stats_reset: device milliseconds of the kernels
 * (CUDA events on the context's stream), kernel launches and calls. Lets a host that drives the stages through the reference's
 * own frontend attribute a whole compress() to kernels. Returns 0 for an unknown id. */
enum
{
	B200_STAT_ETC1S_ENCODE_BLOCKS = 0, B200_STAT_ETC1S_ENDPOINT_CLUSTERS = 1, B200_STAT_ETC1S_REFINE = 2, B200_STAT_ETC1S_DETERMINE_SELECTORS = 3,
	B200_STAT_ETC1S_FIND_SELECTOR_CLUSTERS = 4, B200_STAT_ETC1S_SELECTOR_CODEBOOK = 5, B200_STAT_TSVQ = 6, B200_STAT_UASTC_ENCODE = 7, B200_STAT_UASTC_RDO = 8,
	B200_STAT_ETC1S_REOPTIMIZE_CLUSTERS = 9, B200_STAT_ETC1S_BACKEND_PREDICTION = 10, B200_STAT_COUNT = 11
};
int b200_stats_get(const b200_context* ctx, uint32_t stat_id, float* pKernel_ms, uint32_t* pLaunches, uint32_t* pCalls);
void b200_stats_reset(b200_context* ctx);
/* Same, summed over every context this process has created (a host that never sees the contexts, e.g. basis_compress()). */
int b200_global_stats_get(uint32_t stat_id, float* pKernel_ms, uint32_t* pLaunches, uint32_t* pCalls);
void b200_global_stats_reset(void);

/* Kernel launches issued by this process through any context since load (lets a host that only sees the reference's API,
 * e.g. the drop-in build of INTEGRATION.md section 1, confirm that the GPU path really ran). */
uint64_t b200_global_launch_count(void);

/* Per-stage device time of the last UASTC encode call, summed over its chunks (CUDA events around each kernel on the
 * context's stream): stage 0 = classify/rank, 1 = candidate generation + scoring, 2 = select/hints/pack. */
float b200_last_stage_ms(const b200_context* ctx, uint32_t stage);

/* Event pair on the context's stream for timing a span of calls (e.g. K benchmark steps) on the device. */
int b200_timer_start(b200_context* ctx);
float b200_timer_stop_ms(b200_context* ctx); /* synchronises the stream; < 0 on error */

/* ---- mip generation --------------------------------------------------------------------------------------------------------- */

/* basisu::image_resample for 8-bit images (encoder/basisu_enc.cpp:1022-1171), which basis_compressor::generate_mipmaps calls per
 * level (encoder/basisu_comp.cpp:2146-2230): separable filtering of `num_comps` channels starting at `first_comp` with the
 * reference's contributor lists -- pClist_?_offsets[i] .. [i + 1] index the contributors (source index, float weight) of
 * destination column / row i, exactly Resampler::get_clist_x() / get_clist_y() (encoder/basisu_resampler.h:32-42) flattened --
 * in the reference's operation order, axis order (resampler.cpp:772-806), [0, 1] clamp and 8-bit conversion, so pDst receives the
 * bytes image_resample writes; the other channels of pDst keep their contents. sRGB filtering: pSrgb_to_linear[256] and
 * pLinear_to_srgb[8192] are the two tables image_resample builds (enc.cpp:1061-1075), else both NULL. HOST pointers. */
typedef struct b200_resample_contrib { float weight; uint32_t pixel; } b200_resample_contrib;
int b200_image_resample_rgba8(b200_context* ctx, const void* pSrc, uint32_t src_w, uint32_t src_h, size_t src_pitch_bytes, void* pDst, uint32_t dst_w, uint32_t dst_h,
	size_t dst_pitch_bytes, const uint32_t* pClist_x_offsets, const b200_resample_contrib* pClist_x, const uint32_t* pClist_y_offsets, const b200_resample_contrib* pClist_y,
	uint32_t first_comp, uint32_t num_comps, const float* pSrgb_to_linear, const uint8_t* pLinear_to_srgb);

#ifdef __cplusplus
}
#endif
#endif /* BASISU_B200_H */
