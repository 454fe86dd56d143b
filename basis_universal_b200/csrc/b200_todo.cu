// b200_todo.cu -- entry points declared in include/basisu_b200.h whose kernels are not written yet. They fail loudly
// (return 0 with an error string) so a caller falls back to its own CPU code, exactly like the reference's OpenCL seam.
#include "b200_internal.h"

#define B200_TODO(name) { if (ctx) ctx->fail(name ": not implemented in this build"); return 0; }

extern "C" int b200_uastc_rdo(b200_context* ctx, uint32_t, void*, const void*, const b200_uastc_rdo_params*, uint32_t, uint32_t) B200_TODO("b200_uastc_rdo")
extern "C" int b200_etc1s_set_pixel_blocks(b200_context* ctx, uint32_t, const void*) B200_TODO("b200_etc1s_set_pixel_blocks")
extern "C" int b200_etc1s_encode_blocks(b200_context* ctx, void*, int, uint32_t) B200_TODO("b200_etc1s_encode_blocks")
extern "C" int b200_etc1s_encode_pixel_clusters(b200_context* ctx, void*, uint32_t, const b200_pixel_cluster*, uint64_t, const void*, const uint32_t*, int, uint32_t) B200_TODO("b200_etc1s_encode_pixel_clusters")
extern "C" int b200_etc1s_refine_endpoint_clusterization(b200_context* ctx, const b200_block_info*, uint32_t, const b200_endpoint_cluster*, const uint32_t*, uint32_t*, int) B200_TODO("b200_etc1s_refine_endpoint_clusterization")
extern "C" int b200_etc1s_find_optimal_selector_clusters_for_each_block(b200_context* ctx, const b200_fosc_block*, uint32_t, const b200_fosc_selector*, const uint32_t*, uint32_t*, int) B200_TODO("b200_etc1s_find_optimal_selector_clusters_for_each_block")
extern "C" int b200_etc1s_determine_selectors(b200_context* ctx, const void*, void*, int) B200_TODO("b200_etc1s_determine_selectors")
