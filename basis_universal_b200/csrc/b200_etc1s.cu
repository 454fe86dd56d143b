// b200_etc1s.cu -- CUDA (sm_100a) implementations of the five ETC1S frontend stages behind the reference's GPU seam
// (encoder/basisu_opencl.h:46-141; kernels in bin/ocl_kernels.cl) and their C-ABI entry points (include/basisu_b200.h).
//
//   k_etc1s_encode_blocks       1 thread / block                       opencl_encode_etc1s_blocks          (cl:984)
//   k_etc1s_pixel_clusters      1 WARP / cluster, lanes stride texels  opencl_encode_etc1s_pixel_clusters  (cl:1013)
//   k_etc1s_refine              1 thread / block, cluster-sorted order opencl_refine_endpoint_clusterization (cl:1063)
//   k_etc1s_fosc                1 WARP / block, lanes stride selectors opencl_find_optimal_selector_clusters_for_each_block (cl:1159)
//   k_etc1s_determine_selectors 1 thread / block                       opencl_determine_selectors          (cl:1227)
//
// The reference launches one work-item per block/cluster with fully serial inner loops. Where the serial loop is a sum or an
// arg-min over many items (cluster texels, codebook entries) it is spread over the 32 lanes of a warp here: the sums are
// integer (exact in any order) and the arg-min rules are reproduced as lexicographic (error, index) minima.
#include "b200_internal.h"
#include "bu_etc1s.h"

using namespace bu;

#include "b200_tables.cuh"

__device__ __forceinline__ void load_block16(const uint4* __restrict__ blocks, uint32_t i, uint32_t* px)
{
	const uint4* p = blocks + (size_t)i * 4;
#pragma unroll
	for (int r = 0; r < 4; r++)
	{
		const uint4 v = __ldg(p + r);
		px[r * 4 + 0] = v.x; px[r * 4 + 1] = v.y; px[r * 4 + 2] = v.z; px[r * 4 + 3] = v.w;
	}
}

__global__ void __launch_bounds__(128) k_etc1s_encode_blocks(const uint4* __restrict__ blocks, uint32_t first, uint32_t n, uint64_t* __restrict__ out, int perceptual, uint32_t total_perms, int flavour)
{
	const uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x; // [first, n): this rank's block range
	if (i >= n) return;
	uint32_t px[16];
	load_block16(blocks, i, px);
	out[i] = etc1s_encode_block(&d_tables, perceptual != 0, total_perms, px, flavour);
}

__global__ void __launch_bounds__(128) k_etc1s_determine_selectors(const uint4* __restrict__ blocks, uint32_t first, uint32_t n, const uint32_t* __restrict__ color5_inten, uint64_t* __restrict__ out, int perceptual)
{
	const uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t px[16];
	load_block16(blocks, i, px);
	out[i] = etc1s_determine_selectors(&d_tables, perceptual != 0, px, color5_inten[i]);
}

// ---- pixel clusters: one warp per cluster ---------------------------------------------------------------------------------

__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v)
{
#pragma unroll
	for (int m = 16; m >= 1; m >>= 1)
	{
		uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
		lo = __shfl_xor_sync(0xffffffffu, lo, m); hi = __shfl_xor_sync(0xffffffffu, hi, m);
		v += ((uint64_t)hi << 32) | lo;
	}
	return v;
}

// Where a cluster's texels come from: an explicit (texel, weight) list (the reference's seam), or the cluster's blocks read in
// place from the resident source-block array, every texel with weight 1 (b200_etc1s_encode_endpoint_clusters).
struct px_list_src
{
	static constexpr bool forced = false, unit_weights = false;
	__device__ __forceinline__ uint32_t selector(uint64_t) const { return 0; }
	const uint32_t* __restrict__ px; const uint32_t* __restrict__ wts; uint64_t n;
	__device__ __forceinline__ uint32_t pixel(uint64_t i) const { return px[i]; }
	__device__ __forceinline__ uint32_t weight(uint64_t i) const { return wts[i]; }
};
struct px_blocks_src
{
	static constexpr bool forced = false, unit_weights = true;
	__device__ __forceinline__ uint32_t selector(uint64_t) const { return 0; }
	const uint32_t* __restrict__ blocks; const uint32_t* __restrict__ bidx; uint64_t n; // n = 16 * number of blocks
	__device__ __forceinline__ uint32_t pixel(uint64_t i) const { return __ldg(blocks + (size_t)__ldg(bidx + (i >> 4)) * 16 + (i & 15)); }
	__device__ __forceinline__ uint32_t weight(uint64_t) const { return 1u; }
};

// A cluster is optimised by a TEAM whose members stride its texels: a warp (shuffle reductions), or for the few very large
// clusters a whole CTA (warp reductions + shared memory). All decisions are taken on team-reduced values, so control flow is
// uniform across the team and both teams compute exactly what one CPU thread computes.
struct warp_team
{
	uint32_t rank, bloom_word;
	static constexpr uint32_t size = 32;
	__device__ explicit warp_team() : rank(threadIdx.x & 31), bloom_word(0) {}
	__device__ __forceinline__ uint64_t sum(uint64_t v) { return warp_sum_u64(v); }
	__device__ __forceinline__ void minmax(uint32_t& mn, uint32_t& mx)
	{
#pragma unroll
		for (int m = 16; m >= 1; m >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, m)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, m)); }
	}
	// The CPU optimiser's 1024-bit Bloom filter of tried base colours (etc.cpp:1072), one 32-bit word per lane.
	__device__ __forceinline__ bool bloom_test_and_set(uint32_t r5, uint32_t g5, uint32_t b5)
	{
		const uint32_t kh = hash_hsieh3(r5, g5, b5);
		const uint32_t h0 = kh & 1023, h1 = (kh >> 10) & 1023;
		const uint32_t w0 = __shfl_sync(0xffffffffu, bloom_word, h0 >> 5), w1 = __shfl_sync(0xffffffffu, bloom_word, h1 >> 5);
		if ((w0 >> (h0 & 31)) & (w1 >> (h1 & 31)) & 1) return false;
		if (rank == (h0 >> 5)) bloom_word |= 1u << (h0 & 31);
		if (rank == (h1 >> 5)) bloom_word |= 1u << (h1 & 31);
		return true;
	}
};

#define ETC1S_CTA_TEAM 256
struct cta_team_smem { uint64_t part[ETC1S_CTA_TEAM / 32]; uint32_t mn[ETC1S_CTA_TEAM / 32], mx[ETC1S_CTA_TEAM / 32]; uint32_t bloom[32]; };
struct cta_team
{
	uint32_t rank;
	cta_team_smem* S;
	static constexpr uint32_t size = ETC1S_CTA_TEAM;
	__device__ explicit cta_team(cta_team_smem* s) : rank(threadIdx.x), S(s) { if (threadIdx.x < 32) s->bloom[threadIdx.x] = 0; __syncthreads(); }
	__device__ __forceinline__ uint64_t sum(uint64_t v)
	{
		v = warp_sum_u64(v);
		if (!(rank & 31)) S->part[rank >> 5] = v;
		__syncthreads();
		uint64_t t = 0;
#pragma unroll
		for (int w = 0; w < ETC1S_CTA_TEAM / 32; w++) t += S->part[w];
		__syncthreads();
		return t;
	}
	__device__ __forceinline__ void minmax(uint32_t& mn, uint32_t& mx)
	{
#pragma unroll
		for (int m = 16; m >= 1; m >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, m)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, m)); }
		if (!(rank & 31)) { S->mn[rank >> 5] = mn; S->mx[rank >> 5] = mx; }
		__syncthreads();
#pragma unroll
		for (int w = 0; w < ETC1S_CTA_TEAM / 32; w++) { mn = min(mn, S->mn[w]); mx = max(mx, S->mx[w]); }
		__syncthreads();
	}
	__device__ __forceinline__ bool bloom_test_and_set(uint32_t r5, uint32_t g5, uint32_t b5)
	{
		const uint32_t kh = hash_hsieh3(r5, g5, b5);
		const uint32_t h0 = kh & 1023, h1 = (kh >> 10) & 1023;
		const bool seen = ((S->bloom[h0 >> 5] >> (h0 & 31)) & (S->bloom[h1 >> 5] >> (h1 & 31)) & 1) != 0;
		__syncthreads();
		if (!seen && !rank) { S->bloom[h0 >> 5] |= 1u << (h0 & 31); S->bloom[h1 >> 5] |= 1u << (h1 & 31); }
		__syncthreads();
		return !seen;
	}
};

// The same with an imposed selector per texel (etc1_optimizer::params::m_pForce_selectors, etc.cpp:1137, 1198-1202): `sels[k]` holds
// the 16 selectors of the k-th listed block, texel (x, y) at bits 2 * (x + 4 * y).
struct px_blocks_forced_src
{
	static constexpr bool forced = true, unit_weights = true;
	const uint32_t* __restrict__ blocks; const uint32_t* __restrict__ bidx; const uint32_t* __restrict__ sels; uint64_t n;
	__device__ __forceinline__ uint32_t pixel(uint64_t i) const { return __ldg(blocks + (size_t)__ldg(bidx + (i >> 4)) * 16 + (i & 15)); }
	__device__ __forceinline__ uint32_t weight(uint64_t) const { return 1u; }
	__device__ __forceinline__ uint32_t selector(uint64_t i) const { return (__ldg(sels + (i >> 4)) >> (2 * (uint32_t)(i & 15))) & 3u; }
};

__global__ void __launch_bounds__(128) k_etc1s_pixel_clusters(const b200_pixel_cluster* __restrict__ clusters, uint32_t total_clusters,
	const uint32_t* __restrict__ pixels, const uint32_t* __restrict__ weights, uint64_t* __restrict__ out, int perceptual_i, uint32_t total_perms, int flavour)
{
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (warp >= total_clusters) return;
	px_list_src src;
	src.px = pixels + clusters[warp].first_pixel_index; src.wts = weights + clusters[warp].first_pixel_index; src.n = clusters[warp].total_pixels;
	warp_team team;
	const uint64_t r = cluster_optimize(&d_tables, perceptual_i != 0, src, team, total_perms, flavour);
	if (lane == 0) out[warp] = r;
}

// Endpoint clusters as CSR lists of block indices into the resident source blocks; `order` lists the clusters largest first
// so that the long ones start early (a cluster is one warp's serial work).
__global__ void __launch_bounds__(128) k_etc1s_endpoint_clusters(const uint32_t* __restrict__ blocks, const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ bidx,
	const uint32_t* __restrict__ order, uint32_t first, uint32_t total_clusters, uint64_t* __restrict__ out, int perceptual_i, uint32_t total_perms, int flavour, uint32_t rank, uint32_t world)
{
	const uint32_t warp = first + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * world + rank, lane = threadIdx.x & 31; // every world-th cluster of the size-sorted order, after the big ones
	if (warp >= total_clusters) return;
	const uint32_t c = order[warp];
	px_blocks_src src;
	src.blocks = blocks; src.bidx = bidx + offsets[c]; src.n = (uint64_t)(offsets[c + 1] - offsets[c]) * 16;
	uint64_t r = 0;
	warp_team team;
	if (src.n) r = cluster_optimize(&d_tables, perceptual_i != 0, src, team, total_perms, flavour);
	if (lane == 0) out[c] = r;
}

// The same for the largest clusters (>= ETC1S_BIG_CLUSTER_BLOCKS blocks), one CTA each: `order[0 .. n_big)` are those clusters.
#define ETC1S_BIG_CLUSTER_BLOCKS 128
__global__ void __launch_bounds__(ETC1S_CTA_TEAM) k_etc1s_endpoint_clusters_big(const uint32_t* __restrict__ blocks, const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ bidx,
	const uint32_t* __restrict__ order, uint32_t n_big, uint64_t* __restrict__ out, int perceptual_i, uint32_t total_perms, int flavour, uint32_t rank, uint32_t world)
{
	__shared__ cta_team_smem S;
	const uint32_t k = blockIdx.x * world + rank;
	if (k >= n_big) return;
	const uint32_t c = order[k];
	px_blocks_src src;
	src.blocks = blocks; src.bidx = bidx + offsets[c]; src.n = (uint64_t)(offsets[c + 1] - offsets[c]) * 16;
	cta_team team(&S);
	const uint64_t r = cluster_optimize(&d_tables, perceptual_i != 0, src, team, total_perms, flavour);
	if (!threadIdx.x) out[c] = r;
}

// ---- reoptimize_remapped_endpoints (frontend.cpp:2996-3090): per endpoint cluster, the optimiser with the blocks' selectors imposed,
// plus the error of the cluster's current endpoint with the same selectors. out3[c] = { packed colour + table, new error, current error }.
struct reopt_out { uint64_t packed, new_err, cur_err; };
template<typename Team> __device__ void reoptimize_cluster(const bu_tables* T, bool perceptual, const uint32_t* __restrict__ blocks, const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ bidx,
	const uint32_t* __restrict__ sels, const uint32_t* __restrict__ cur_c5i, uint32_t c, Team& team, uint32_t total_perms, int flavour, reopt_out* out, bool writer)
{
	reopt_out r; r.packed = 0; r.new_err = 0; r.cur_err = 0;
	const uint32_t cur = cur_c5i[c];
	const uint64_t n = (uint64_t)(offsets[c + 1] - offsets[c]) * 16;
	if (n && sels)
	{
		px_blocks_forced_src src;
		src.blocks = blocks; src.bidx = bidx + offsets[c]; src.sels = sels + offsets[c]; src.n = n;
		r.cur_err = cluster_endpoint_error(T, perceptual, src, team, cur & 255, (cur >> 8) & 255, (cur >> 16) & 255, cur >> 24);
		r.packed = cluster_optimize(T, perceptual, src, team, total_perms, flavour, &r.new_err);
	}
	else if (n)
	{
		// selectors free (generate_endpoint_codebook at step >= 1)
		px_blocks_src src;
		src.blocks = blocks; src.bidx = bidx + offsets[c]; src.n = n;
		r.cur_err = cluster_endpoint_error(T, perceptual, src, team, cur & 255, (cur >> 8) & 255, (cur >> 16) & 255, cur >> 24);
		r.packed = cluster_optimize(T, perceptual, src, team, total_perms, flavour, &r.new_err);
	}
	if (writer) out[c] = r;
}
__global__ void __launch_bounds__(128) k_etc1s_reoptimize_clusters(const uint32_t* __restrict__ blocks, const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ bidx, const uint32_t* __restrict__ sels,
	const uint32_t* __restrict__ cur_c5i, const uint32_t* __restrict__ order, uint32_t first, uint32_t total_clusters, reopt_out* __restrict__ out, int perceptual_i, uint32_t total_perms, int flavour, uint32_t rank, uint32_t world)
{
	const uint32_t warp = first + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * world + rank;
	if (warp >= total_clusters) return;
	warp_team team;
	reoptimize_cluster(&d_tables, perceptual_i != 0, blocks, offsets, bidx, sels, cur_c5i, order[warp], team, total_perms, flavour, out, (threadIdx.x & 31) == 0);
}
__global__ void __launch_bounds__(ETC1S_CTA_TEAM) k_etc1s_reoptimize_clusters_big(const uint32_t* __restrict__ blocks, const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ bidx, const uint32_t* __restrict__ sels,
	const uint32_t* __restrict__ cur_c5i, const uint32_t* __restrict__ order, uint32_t n_big, reopt_out* __restrict__ out, int perceptual_i, uint32_t total_perms, int flavour, uint32_t rank, uint32_t world)
{
	__shared__ cta_team_smem S;
	const uint32_t k = blockIdx.x * world + rank;
	if (k >= n_big) return;
	cta_team team(&S);
	reoptimize_cluster(&d_tables, perceptual_i != 0, blocks, offsets, bidx, sels, cur_c5i, order[k], team, total_perms, flavour, out, threadIdx.x == 0);
}

// ---- compute_endpoint_subblock_error_vec (frontend.cpp:1006-1082): per block, the error of each of its two subblocks (texels 0-7 and
// 8-15 of the flipped layout) against the endpoint of the block's cluster, best of the four colours per texel.
__global__ void __launch_bounds__(128) k_etc1s_subblock_errors(const uint4* __restrict__ blocks, const uint32_t* __restrict__ c5i, uint32_t first, uint32_t last, uint2* __restrict__ out, int perceptual_i)
{
	const uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= last) return;
	const bool perceptual = perceptual_i != 0;
	const uint32_t e = c5i[i];
	// Reference behaviour kept: frontend.cpp:1043 passes the UNSCALED 5-bit colour to get_block_colors5(..., scaled = true), so the four
	// colours are (r5 + m, g5 + m, b5 + m) clamped, not the expanded 8-bit base plus m. The ranking of subblocks that
	// introduce_new_endpoint_clusters works from is defined by that.
	uint32_t colors[4];
	for (int k = 0; k < 4; k++)
	{
		const int m = d_tables.etc1_inten[(e >> 24) * 4 + k];
		colors[k] = px_make((uint32_t)clamp255i((int)(e & 255) + m), (uint32_t)clamp255i((int)((e >> 8) & 255) + m), (uint32_t)clamp255i((int)((e >> 16) & 255) + m), 255);
	}
	uint32_t err[2] = { 0, 0 }; // <= 8 texels * max distance: fits 32 bits for both metrics
	for (int q = 0; q < 4; q++)
	{
		const uint4 v = __ldg(blocks + (size_t)i * 4 + q);
		const uint32_t px[4] = { v.x, v.y, v.z, v.w };
		for (int k = 0; k < 4; k++)
		{
			uint32_t be = etc_color_distance(perceptual, px[k], colors[0]);
			be = min(be, etc_color_distance(perceptual, px[k], colors[1]));
			be = min(be, etc_color_distance(perceptual, px[k], colors[2]));
			be = min(be, etc_color_distance(perceptual, px[k], colors[3]));
			err[q >> 1] += be;
		}
	}
	out[i] = make_uint2(err[0], err[1]);
}

// ---- create_optimized_selector_codebook: one warp per selector cluster ---------------------------------------------------------
// Lane = (half, texel): the two half-warps take alternate member blocks; each lane accumulates, for its texel, the error of the
// four block colours of every member block (decoded from the block's own endpoint) against the source texel.
__global__ void __launch_bounds__(128) k_etc1s_selector_codebook(const uint32_t* __restrict__ blocks, const uint2* __restrict__ etc_blocks, const uint32_t* __restrict__ offsets,
	const uint32_t* __restrict__ bidx, uint32_t total_clusters, uint32_t* __restrict__ out, int perceptual_i, uint32_t rank, uint32_t world)
{
	const uint32_t warp = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * world + rank, lane = threadIdx.x & 31;
	if (warp >= total_clusters) return;
	const bool perceptual = perceptual_i != 0;
	const uint32_t first = offsets[warp], last = offsets[warp + 1];
	const uint32_t texel = lane & 15; // x + 4 * y
	uint64_t err[4] = { 0, 0, 0, 0 };
	for (uint32_t k = first + (lane >> 4); k < last; k += 2)
	{
		const uint32_t bi = __ldg(bidx + k);
		const uint32_t w = __ldg(&etc_blocks[bi]).x; // bytes 0..3: R5<<3 | dR3, G5<<3 | dG3, B5<<3 | dB3, inten0<<5 | inten1<<2 | diff | flip
		uint32_t colors[4];
		etc1s_block_colors(&d_tables, (w >> 3) & 31, (w >> 11) & 31, (w >> 19) & 31, (w >> 29) & 7, colors); // get_block_colors(., 0): subblock 0's base + table
		const uint32_t p = __ldg(blocks + (size_t)bi * 16 + texel);
#pragma unroll
		for (int s = 0; s < 4; s++) err[s] += etc_color_distance(perceptual, colors[s], p);
	}
	uint32_t best = 0;
	uint64_t best_err = 0;
#pragma unroll
	for (int s = 0; s < 4; s++)
	{
		uint32_t lo = (uint32_t)err[s], hi = (uint32_t)(err[s] >> 32);
		lo = __shfl_xor_sync(0xffffffffu, lo, 16); hi = __shfl_xor_sync(0xffffffffu, hi, 16);
		const uint64_t e = err[s] + (((uint64_t)hi << 32) | lo);
		if (s == 0 || e < best_err) { best_err = e; best = (uint32_t)s; } // first strictly smaller (frontend.cpp:2326-2336)
	}
	uint32_t packed = (lane < 16) ? (best << (texel * 2)) : 0u;
#pragma unroll
	for (int m = 8; m >= 1; m >>= 1) packed |= __shfl_xor_sync(0xffffffffu, packed, m);
	if (lane == 0) out[warp] = (first == last) ? 0u : packed;
}

// ---- refine_endpoint_clusterization ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(128) k_etc1s_refine(const uint4* __restrict__ blocks, uint32_t first, uint32_t n, const b200_block_info* __restrict__ info,
	const b200_endpoint_cluster* __restrict__ clusters, const uint32_t* __restrict__ sorted, uint32_t* __restrict__ out, int perceptual_i)
{
	const uint32_t gid = first + blockIdx.x * blockDim.x + threadIdx.x; // position in the cluster-sorted order
	if (gid >= n) return;
	const uint32_t bi = sorted[gid];
	const bool perceptual = perceptual_i != 0;
	uint32_t px[16];
	load_block16(blocks, bi, px);
	const b200_block_info in = info[bi];
	uint64_t best_err = UINT64_MAX;
	uint32_t best_index = 0;
	for (uint32_t k = 0; k < in.num_clusters; k++)
	{
		const b200_endpoint_cluster c = clusters[(uint32_t)in.first_cluster_ofs + k];
		if (c.etc_inten > in.cur_cluster_etc_inten) continue;
		const uint64_t e = etc1s_block_error(&d_tables, perceptual, px, c.unscaled_r, c.unscaled_g, c.unscaled_b, c.etc_inten);
		if (e < best_err || (c.cluster_index == in.cur_cluster_index && e == best_err))
		{
			best_err = e;
			best_index = c.cluster_index;
			if (!best_err) break;
		}
	}
	out[bi] = best_index;
}

// ---- find_optimal_selector_clusters_for_each_block: one warp per block -------------------------------------------------------

__global__ void __launch_bounds__(128) k_etc1s_fosc(const uint4* __restrict__ blocks, uint32_t first, uint32_t n, const b200_fosc_block* __restrict__ info,
	const uint32_t* __restrict__ selectors, const uint32_t* __restrict__ cluster_indices, uint32_t* __restrict__ out, int perceptual_i)
{
	__shared__ uint32_t s_err[4][64]; // per warp: [selector 0..3][texel] errors
	const uint32_t warp_in_cta = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t bi = first + blockIdx.x * (blockDim.x >> 5) + warp_in_cta;
	if (bi >= n) return;
	const bool perceptual = perceptual_i != 0;
	const b200_fosc_block in = info[bi];
	uint32_t colors[4];
	etc1s_block_colors(&d_tables, in.etc_r, in.etc_g, in.etc_b, in.etc_a, colors);
	// 64 table entries, two per lane: entry e = sel * 16 + texel
	const uint32_t* bp = reinterpret_cast<const uint32_t*>(blocks) + (size_t)bi * 16;
	for (uint32_t e = lane; e < 64; e += 32)
		s_err[warp_in_cta][e] = etc_color_distance(perceptual, __ldg(bp + (e & 15)), colors[e >> 4]);
	__syncwarp();

	const uint32_t* tab = s_err[warp_in_cta];
	const uint32_t* sel = selectors + in.first_selector;
	uint64_t best_err = UINT64_MAX;
	uint32_t best_index = 0xFFFFFFFFu;
	for (uint32_t k = lane; k < in.num_selectors; k += 32)
	{
		uint32_t s = __ldg(sel + k);
		uint64_t total = 0;
#pragma unroll
		for (int i = 0; i < 16; i++, s >>= 2) total += tab[(s & 3) * 16 + i];
		if (total < best_err) { best_err = total; best_index = k; } // ascending k per lane: strict < keeps the earliest
	}
#pragma unroll
	for (int m = 16; m >= 1; m >>= 1)
	{
		uint32_t lo = (uint32_t)best_err, hi = (uint32_t)(best_err >> 32);
		lo = __shfl_xor_sync(0xffffffffu, lo, m); hi = __shfl_xor_sync(0xffffffffu, hi, m);
		const uint64_t oe = ((uint64_t)hi << 32) | lo;
		const uint32_t oi = __shfl_xor_sync(0xffffffffu, best_index, m);
		if (oe < best_err || (oe == best_err && oi < best_index)) { best_err = oe; best_index = oi; }
	}
	if (lane == 0)
	{
		if (best_index == 0xFFFFFFFFu) best_index = 0; // num_selectors == 0: the kernel reads entry 0 (cl:1216)
		out[bi] = cluster_indices[in.first_selector + best_index];
	}
}

// ---- endpoint training-set histogram (multi-GPU exchange point) ------------------------------------------------------------------
// A block's endpoint training vector is a function of its 18-bit (r5, g5, b5, inten) key only (init_endpoint_training_vectors,
// frontend.cpp:843-857: low/high block colours / 255, inserted twice with weight 1), and generate_hierarchical_codebook_threaded
// starts by merging identical vectors into (vector, weight) pairs in lexicographic order (encoder/basisu_enc.h:2228-2260).
// A dense u32[2^18] count per key is therefore a complete, order-free description of a shard's training set: ranks SUM
// all-reduce it (1 MiB over NVLink) and every rank holds the global weighted unique set.
__global__ void __launch_bounds__(256) k_etc1s_endpoint_histogram(const uint2* __restrict__ etc_blocks, uint32_t n, uint32_t* __restrict__ hist)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t w = __ldg(&etc_blocks[i]).x; // bytes 0..3: R5<<3, G5<<3, B5<<3, inten0<<5 | inten1<<2 | diff | flip
	const uint32_t key = (((w >> 3) & 31) << 13) | (((w >> 11) & 31) << 8) | (((w >> 19) & 31) << 3) | ((w >> 29) & 7);
	atomicAdd(hist + key, 2u); // two subblocks per block, weight 1 each
}

// ---- selector training set (second multi-GPU exchange point) ----------------------------------------------------------------
// Per block: the 16 selectors as a 32-bit key (texel x + 4y at bits 2(x + 4y)) and the weight generate_selector_clusters gives
// the block's training vector (frontend.cpp:2156-2179): colour distance between the low and high block colours of the
// sub-block with the larger intensity table, / 300, clamped to 1..4096. Identical keys are merged by the clusterer
// (enc.h:2228-2260), so ranks exchange (key, summed weight) pairs.
__global__ void __launch_bounds__(256) k_etc1s_selector_training(const uint2* __restrict__ etc_blocks, uint32_t n, int perceptual, uint32_t* __restrict__ keys, uint32_t* __restrict__ weights)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint2 w = __ldg(&etc_blocks[i]);
	const uint32_t b0 = w.x & 255, b1 = (w.x >> 8) & 255, b2 = (w.x >> 16) & 255, b3 = w.x >> 24;
	const uint32_t msb = ((w.y & 255) << 8) | ((w.y >> 8) & 255), lsb = (((w.y >> 16) & 255) << 8) | (w.y >> 24);
	uint32_t key = 0;
	for (uint32_t y = 0; y < 4; y++)
		for (uint32_t x = 0; x < 4; x++)
		{
			const uint32_t bit = x * 4 + y;
			const uint32_t raw = (((msb >> bit) & 1) << 1) | ((lsb >> bit) & 1);
			const uint32_t sel = (0x1Eu >> (raw * 2)) & 3u; // g_etc1_to_selector_index = { 2, 3, 1, 0 }
			key |= sel << ((x + y * 4) * 2);
		}
	const uint32_t inten0 = b3 >> 5, inten1 = (b3 >> 2) & 7;
	const uint32_t sub = (inten0 > inten1) ? 0u : 1u;
	int c5[3] = { (int)(b0 >> 3), (int)(b1 >> 3), (int)(b2 >> 3) };
	if (sub)
	{
		// differential mode: second sub-block = base5 + signed 3-bit delta (always 0 for ETC1S blocks), clamped as unpack_color5 does
		const uint32_t d3[3] = { b0 & 7, b1 & 7, b2 & 7 };
		for (int c = 0; c < 3; c++) c5[c] = clampi(c5[c] + (int)(d3[c] >= 4 ? d3[c] - 8 : d3[c]), 0, 31);
	}
	uint32_t colors[4];
	etc1s_block_colors(&d_tables, (uint32_t)c5[0], (uint32_t)c5[1], (uint32_t)c5[2], sub ? inten1 : inten0, colors);
	const uint32_t dist = etc_color_distance(perceptual != 0, colors[0], colors[3]);
	keys[i] = key;
	weights[i] = min(max(dist / 300u, 1u), 4096u);
}

// ---- C ABI ------------------------------------------------------------------------------------------------------------------

#define ETC_CHECK_BLOCKS(ctx, name) do { if (!(ctx)) return 0; if (!(ctx)->activate()) return 0; \
	if (!(ctx)->d_etc_blocks || !(ctx)->etc_total_blocks) { (ctx)->fail(name ": b200_etc1s_set_pixel_blocks has not been called"); return 0; } } while (0)

static bool upload(b200_context* ctx, int slot, const void* host, size_t bytes)
{
	if (!ctx->reserve(ctx->d_aux[slot], ctx->aux_cap[slot], bytes ? bytes : 16)) return false;
	if (bytes)
	{
		const cudaError_t e = cudaMemcpyAsync(ctx->d_aux[slot], host, bytes, cudaMemcpyHostToDevice, ctx->stream);
		if (e != cudaSuccess) { ctx->fail_cuda("cudaMemcpyAsync(H2D)", e); return false; }
	}
	return true;
}

// This rank's share [first, last) of n per-block units: contiguous ranges, i.e. block rows.
static void shard_range(const b200_context* ctx, uint32_t n, uint32_t& first, uint32_t& last)
{
	b200_shard_range(n, (uint32_t)ctx->rank, (uint32_t)(ctx->world > 1 ? ctx->world : 1), &first, &last);
}
// With several ranks the output buffer is zeroed before the kernel so that the all-reduce in finish() is a merge.
static bool shard_prepare_output(b200_context* ctx, int slot, size_t bytes)
{
	if (ctx->world <= 1) return true;
	const cudaError_t e = cudaMemsetAsync(ctx->d_aux[slot], 0, bytes, ctx->stream);
	if (e != cudaSuccess) { ctx->fail_cuda("cudaMemsetAsync", e); return false; }
	return true;
}

static int finish(b200_context* ctx, void* host_out, int slot, size_t bytes, int stat_id)
{
	B200_CUDA_OK(ctx, cudaGetLastError());
	B200_CUDA_OK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
	if (!b200_merge_u32(ctx, ctx->d_aux[slot], bytes / 4)) return 0; // every output element is 4 or 8 bytes
	B200_CUDA_OK(ctx, cudaMemcpyAsync(host_out, ctx->d_aux[slot], bytes, cudaMemcpyDeviceToHost, ctx->stream));
	B200_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
	B200_CUDA_OK(ctx, cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
	ctx->account(stat_id);
	return 1;
}

extern "C" int b200_etc1s_set_flavour(b200_context* ctx, int flavour)
{
	if (!ctx) return 0;
	if (flavour != B200_ETC1S_FLAVOUR_OPENCL_KERNELS && flavour != B200_ETC1S_FLAVOUR_CPU_OPTIMIZER) { ctx->fail("b200_etc1s_set_flavour: unknown flavour"); return 0; }
	ctx->etc_flavour = flavour;
	return 1;
}

extern "C" int b200_etc1s_set_pixel_blocks(b200_context* ctx, uint32_t total_blocks, const void* pPixel_blocks)
{
	if (!ctx) return 0;
	if (!ctx->activate()) return 0;
	if (!total_blocks || !pPixel_blocks) { ctx->fail("b200_etc1s_set_pixel_blocks: empty input"); return 0; }
	if (!ctx->reserve(ctx->d_etc_blocks, ctx->etc_blocks_cap, (size_t)total_blocks * 64)) return 0;
	B200_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_etc_blocks, pPixel_blocks, (size_t)total_blocks * 64, cudaMemcpyHostToDevice, ctx->stream));
	B200_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
	ctx->etc_total_blocks = total_blocks;
	return 1;
}

extern "C" int b200_etc1s_encode_blocks(b200_context* ctx, void* pOutput_blocks, int perceptual, uint32_t total_perms)
{
	ETC_CHECK_BLOCKS(ctx, "b200_etc1s_encode_blocks");
	if (total_perms > 165) { ctx->fail("b200_etc1s_encode_blocks: total_perms > 165"); return 0; }
	const uint32_t n = ctx->etc_total_blocks;
	if (!ctx->reserve(ctx->d_aux[0], ctx->aux_cap[0], (size_t)n * 8)) return 0;
	ctx->launches = 1; __atomic_add_fetch(&g_b200_total_launches, 1, __ATOMIC_RELAXED);
	uint32_t first, last;
	shard_range(ctx, n, first, last);
	if (!shard_prepare_output(ctx, 0, (size_t)n * 8)) return 0;
	B200_CUDA_OK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
	if (last > first)
		k_etc1s_encode_blocks<<<(last - first + 127) / 128, 128, 0, ctx->stream>>>(static_cast<const uint4*>(ctx->d_etc_blocks), first, last, static_cast<uint64_t*>(ctx->d_aux[0]), perceptual, total_perms, ctx->etc_flavour);
	return finish(ctx, pOutput_blocks, 0, (size_t)n * 8, B200_STAT_ETC1S_ENCODE_BLOCKS);
}

extern "C" int b200_etc1s_encode_pixel_clusters(b200_context* ctx, void* pOutput_blocks, uint32_t total_clusters, const b200_pixel_cluster* pClusters,
	uint64_t total_pixels, const void* pPixels, const uint32_t* pPixel_weights, int perceptual, uint32_t total_perms)
{
	if (!ctx) return 0;
	if (!ctx->activate()) return 0;
	if (!total_clusters) return 1;
	if (total_perms > 165) { ctx->fail("b200_etc1s_encode_pixel_clusters: total_perms > 165"); return 0; }
	if (!upload(ctx, 1, pClusters, (size_t)total_clusters * sizeof(b200_pixel_cluster))) return 0;
	if (!upload(ctx, 2, pPixels, (size_t)total_pixels * 4)) return 0;
	if (!upload(ctx, 3, pPixel_weights, (size_t)total_pixels * 4)) return 0;
	if (!ctx->reserve(ctx->d_aux[0], ctx->aux_cap[0], (size_t)total_clusters * 8)) return 0;
	ctx->launches = 1; __atomic_add_fetch(&g_b200_total_launches, 1, __ATOMIC_RELAXED);
	B200_CUDA_OK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
	const uint32_t warps_per_cta = 4;
	k_etc1s_pixel_clusters<<<(total_clusters + warps_per_cta - 1) / warps_per_cta, warps_per_cta * 32, 0, ctx->stream>>>(
		static_cast<const b200_pixel_cluster*>(ctx->d_aux[1]), total_clusters, static_cast<const uint32_t*>(ctx->d_aux[2]),
		static_cast<const uint32_t*>(ctx->d_aux[3]), static_cast<uint64_t*>(ctx->d_aux[0]), perceptual, total_perms, ctx->etc_flavour);
	return finish(ctx, pOutput_blocks, 0, (size_t)total_clusters * 8, B200_STAT_ETC1S_ENDPOINT_CLUSTERS);
}

extern "C" int b200_etc1s_refine_endpoint_clusterization(b200_context* ctx, const b200_block_info* pPixel_block_info, uint32_t total_clusters,
	const b200_endpoint_cluster* pCluster_info, const uint32_t* pSorted_block_indices, uint32_t* pOutput_cluster_indices, int perceptual)
{
	ETC_CHECK_BLOCKS(ctx, "b200_etc1s_refine_endpoint_clusterization");
	const uint32_t n = ctx->etc_total_blocks;
	if (!upload(ctx, 1, pPixel_block_info, (size_t)n * sizeof(b200_block_info))) return 0;
	if (!upload(ctx, 2, pCluster_info, (size_t)total_clusters * sizeof(b200_endpoint_cluster))) return 0;
	if (!upload(ctx, 3, pSorted_block_indices, (size_t)n * 4)) return 0;
	if (!ctx->reserve(ctx->d_aux[0], ctx->aux_cap[0], (size_t)n * 4)) return 0;
	ctx->launches = 1; __atomic_add_fetch(&g_b200_total_launches, 1, __ATOMIC_RELAXED);
	uint32_t first, last;
	shard_range(ctx, n, first, last);
	if (!shard_prepare_output(ctx, 0, (size_t)n * 4)) return 0;
	B200_CUDA_OK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
	if (last > first)
	k_etc1s_refine<<<(last - first + 127) / 128, 128, 0, ctx->stream>>>(static_cast<const uint4*>(ctx->d_etc_blocks), first, last, static_cast<const b200_block_info*>(ctx->d_aux[1]),
		static_cast<const b200_endpoint_cluster*>(ctx->d_aux[2]), static_cast<const uint32_t*>(ctx->d_aux[3]), static_cast<uint32_t*>(ctx->d_aux[0]), perceptual);
	return finish(ctx, pOutput_cluster_indices, 0, (size_t)n * 4, B200_STAT_ETC1S_REFINE);
}

extern "C" int b200_etc1s_find_optimal_selector_clusters_for_each_block(b200_context* ctx, const b200_fosc_block* pInput_block_info, uint32_t total_input_selectors,
	const b200_fosc_selector* pInput_selectors, const uint32_t* pSelector_cluster_indices, uint32_t* pOutput_selector_cluster_indices, int perceptual)
{
	ETC_CHECK_BLOCKS(ctx, "b200_etc1s_find_optimal_selector_clusters_for_each_block");
	const uint32_t n = ctx->etc_total_blocks;
	if (!upload(ctx, 1, pInput_block_info, (size_t)n * sizeof(b200_fosc_block))) return 0;
	if (!upload(ctx, 2, pInput_selectors, (size_t)total_input_selectors * 4)) return 0;
	if (!upload(ctx, 3, pSelector_cluster_indices, (size_t)total_input_selectors * 4)) return 0;
	if (!ctx->reserve(ctx->d_aux[0], ctx->aux_cap[0], (size_t)n * 4)) return 0;
	ctx->launches = 1; __atomic_add_fetch(&g_b200_total_launches, 1, __ATOMIC_RELAXED);
	uint32_t first, last;
	shard_range(ctx, n, first, last);
	if (!shard_prepare_output(ctx, 0, (size_t)n * 4)) return 0;
	B200_CUDA_OK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
	const uint32_t warps_per_cta = 4;
	if (last > first)
	k_etc1s_fosc<<<(last - first + warps_per_cta - 1) / warps_per_cta, warps_per_cta * 32, 0, ctx->stream>>>(static_cast<const uint4*>(ctx->d_etc_blocks), first, last,
		static_cast<const b200_fosc_block*>(ctx->d_aux[1]), static_cast<const uint32_t*>(ctx->d_aux[2]), static_cast<const uint32_t*>(ctx->d_aux[3]),
		static_cast<uint32_t*>(ctx->d_aux[0]), perceptual);
	return finish(ctx, pOutput_selector_cluster_indices, 0, (size_t)n * 4, B200_STAT_ETC1S_FIND_SELECTOR_CLUSTERS);
}

extern "C" int b200_etc1s_determine_selectors(b200_context* ctx, const void* pInput_etc_color5_and_inten, void* pOutput_blocks, int perceptual)
{
	ETC_CHECK_BLOCKS(ctx, "b200_etc1s_determine_selectors");
	const uint32_t n = ctx->etc_total_blocks;
	if (!upload(ctx, 1, pInput_etc_color5_and_inten, (size_t)n * 4)) return 0;
	if (!ctx->reserve(ctx->d_aux[0], ctx->aux_cap[0], (size_t)n * 8)) return 0;
	ctx->launches = 1; __atomic_add_fetch(&g_b200_total_launches, 1, __ATOMIC_RELAXED);
	uint32_t first, last;
	shard_range(ctx, n, first, last);
	if (!shard_prepare_output(ctx, 0, (size_t)n * 8)) return 0;
	B200_CUDA_OK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
	if (last > first)
	k_etc1s_determine_selectors<<<(last - first + 127) / 128, 128, 0, ctx->stream>>>(static_cast<const uint4*>(ctx->d_etc_blocks), first, last, static_cast<const uint32_t*>(ctx->d_aux[1]),
		static_cast<uint64_t*>(ctx->d_aux[0]), perceptual);
	return finish(ctx, pOutput_blocks, 0, (size_t)n * 8, B200_STAT_ETC1S_DETERMINE_SELECTORS);
}

extern "C" int b200_etc1s_endpoint_histogram_device(b200_context* ctx, const void* dEtc_blocks, uint32_t num_blocks, uint32_t* dHist)
{
	if (!ctx) return 0;
	if (!ctx->activate()) return 0;
	if (!num_blocks) return 1;
	ctx->launches = 1; __atomic_add_fetch(&g_b200_total_launches, 1, __ATOMIC_RELAXED);
	B200_CUDA_OK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
	k_etc1s_endpoint_histogram<<<(num_blocks + 255) / 256, 256, 0, ctx->stream>>>(static_cast<const uint2*>(dEtc_blocks), num_blocks, dHist);
	B200_CUDA_OK(ctx, cudaGetLastError());
	B200_CUDA_OK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
	B200_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
	B200_CUDA_OK(ctx, cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
	return 1;
}

extern "C" int b200_etc1s_endpoint_histogram(b200_context* ctx, const void* pEtc_blocks, uint32_t num_blocks, uint32_t* pHist)
{
	if (!ctx) return 0;
	if (!ctx->activate()) return 0;
	if (!pHist) { ctx->fail("b200_etc1s_endpoint_histogram: null histogram"); return 0; }
	const size_t hist_bytes = sizeof(uint32_t) << 18;
	if (!upload(ctx, 1, pEtc_blocks, (size_t)num_blocks * 8)) return 0;
	if (!ctx->reserve(ctx->d_aux[0], ctx->aux_cap[0], hist_bytes)) return 0;
	B200_CUDA_OK(ctx, cudaMemsetAsync(ctx->d_aux[0], 0, hist_bytes, ctx->stream));
	B200_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
	if (!b200_etc1s_endpoint_histogram_device(ctx, ctx->d_aux[1], num_blocks, static_cast<uint32_t*>(ctx->d_aux[0]))) return 0;
	B200_CUDA_OK(ctx, cudaMemcpy(pHist, ctx->d_aux[0], hist_bytes, cudaMemcpyDeviceToHost));
	return 1;
}

extern "C" int b200_etc1s_selector_training_device(b200_context* ctx, const void* dEtc_blocks, uint32_t num_blocks, int perceptual, uint32_t* dKeys, uint32_t* dWeights)
{
	if (!ctx) return 0;
	if (!ctx->activate()) return 0;
	if (!num_blocks) return 1;
	ctx->launches = 1; __atomic_add_fetch(&g_b200_total_launches, 1, __ATOMIC_RELAXED);
	B200_CUDA_OK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
	k_etc1s_selector_training<<<(num_blocks + 255) / 256, 256, 0, ctx->stream>>>(static_cast<const uint2*>(dEtc_blocks), num_blocks, perceptual, dKeys, dWeights);
	B200_CUDA_OK(ctx, cudaGetLastError());
	B200_CUDA_OK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
	B200_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
	B200_CUDA_OK(ctx, cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
	return 1;
}

extern "C" int b200_etc1s_selector_training(b200_context* ctx, const void* pEtc_blocks, uint32_t num_blocks, int perceptual, uint32_t* pKeys, uint32_t* pWeights)
{
	if (!ctx) return 0;
	if (!ctx->activate()) return 0;
	if (!num_blocks) return 1;
	if (!pEtc_blocks || !pKeys || !pWeights) { ctx->fail("b200_etc1s_selector_training: null buffer"); return 0; }
	if (!upload(ctx, 1, pEtc_blocks, (size_t)num_blocks * 8)) return 0;
	if (!ctx->reserve(ctx->d_aux[0], ctx->aux_cap[0], (size_t)num_blocks * 8)) return 0;
	uint32_t* dk = static_cast<uint32_t*>(ctx->d_aux[0]);
	if (!b200_etc1s_selector_training_device(ctx, ctx->d_aux[1], num_blocks, perceptual, dk, dk + num_blocks)) return 0;
	B200_CUDA_OK(ctx, cudaMemcpy(pKeys, dk, (size_t)num_blocks * 4, cudaMemcpyDeviceToHost));
	B200_CUDA_OK(ctx, cudaMemcpy(pWeights, dk + num_blocks, (size_t)num_blocks * 4, cudaMemcpyDeviceToHost));
	return 1;
}

// Launch order of per-cluster kernels: clusters by size class (floor(log2(size))), largest class first; clusters of >=
// ETC1S_BIG_CLUSTER_BLOCKS blocks (a power of two, so whole classes) form the prefix that gets one CTA each. malloc'd; caller frees.
static uint32_t* cluster_launch_order(const uint32_t* offsets, uint32_t total_clusters, uint32_t& n_big)
{
	uint32_t* order = static_cast<uint32_t*>(malloc((size_t)total_clusters * 4));
	if (!order) return nullptr;
	n_big = 0;
	uint32_t counts[33] = { 0 }, ofs[33];
	auto cls = [&](uint32_t c) { const uint32_t sz = offsets[c + 1] - offsets[c]; uint32_t b = 0; while (b < 32 && (1u << b) <= sz) b++; return 32 - b; };
	for (uint32_t i = 0; i < total_clusters; i++) { counts[cls(i)]++; if (offsets[i + 1] - offsets[i] >= ETC1S_BIG_CLUSTER_BLOCKS) n_big++; }
	uint32_t acc = 0;
	for (int b = 0; b < 33; b++) { ofs[b] = acc; acc += counts[b]; }
	for (uint32_t i = 0; i < total_clusters; i++) order[ofs[cls(i)]++] = i;
	return order;
}

extern "C" int b200_etc1s_encode_endpoint_clusters(b200_context* ctx, void* pOutput_blocks, uint32_t total_clusters, const uint32_t* pCluster_offsets,
	const uint32_t* pCluster_block_indices, int perceptual, uint32_t total_perms)
{
	ETC_CHECK_BLOCKS(ctx, "b200_etc1s_encode_endpoint_clusters");
	if (!total_clusters) return 1;
	if (total_perms > 165) { ctx->fail("b200_etc1s_encode_endpoint_clusters: total_perms > 165"); return 0; }
	if (!pOutput_blocks || !pCluster_offsets || !pCluster_block_indices) { ctx->fail("b200_etc1s_encode_endpoint_clusters: null buffer"); return 0; }
	const uint32_t total_indices = pCluster_offsets[total_clusters];
	for (uint32_t i = 0; i < total_indices; i++)
		if (pCluster_block_indices[i] >= ctx->etc_total_blocks) { ctx->fail("b200_etc1s_encode_endpoint_clusters: block index out of range"); return 0; }
	uint32_t n_big = 0;
	uint32_t* order = cluster_launch_order(pCluster_offsets, total_clusters, n_big);
	if (!order) { ctx->fail("b200_etc1s_encode_endpoint_clusters: out of host memory"); return 0; }
	bool ok = upload(ctx, 1, pCluster_offsets, ((size_t)total_clusters + 1) * 4) && upload(ctx, 2, pCluster_block_indices, (size_t)total_indices * 4) && upload(ctx, 3, order, (size_t)total_clusters * 4);
	if (ok) ok = ctx->reserve(ctx->d_aux[0], ctx->aux_cap[0], (size_t)total_clusters * 8);
	if (ok && cudaStreamSynchronize(ctx->stream) != cudaSuccess) ok = false; // `order` is pageable host memory about to be freed
	free(order);
	if (!ok) return 0;
	ctx->launches = 1; __atomic_add_fetch(&g_b200_total_launches, 1, __ATOMIC_RELAXED);
	if (!shard_prepare_output(ctx, 0, (size_t)total_clusters * 8)) return 0;
	B200_CUDA_OK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
	const uint32_t warps_per_cta = 4, world = (uint32_t)(ctx->world > 1 ? ctx->world : 1);
	if (n_big)
	{
		k_etc1s_endpoint_clusters_big<<<(n_big + world - 1) / world, ETC1S_CTA_TEAM, 0, ctx->stream>>>(static_cast<const uint32_t*>(ctx->d_etc_blocks),
			static_cast<const uint32_t*>(ctx->d_aux[1]), static_cast<const uint32_t*>(ctx->d_aux[2]), static_cast<const uint32_t*>(ctx->d_aux[3]), n_big,
			static_cast<uint64_t*>(ctx->d_aux[0]), perceptual, total_perms, ctx->etc_flavour, (uint32_t)ctx->rank, world);
		ctx->launches++; __atomic_add_fetch(&g_b200_total_launches, 1, __ATOMIC_RELAXED);
	}
	const uint32_t my_clusters = (total_clusters - n_big + world - 1) / world;
	if (my_clusters)
	k_etc1s_endpoint_clusters<<<(my_clusters + warps_per_cta - 1) / warps_per_cta, warps_per_cta * 32, 0, ctx->stream>>>(static_cast<const uint32_t*>(ctx->d_etc_blocks),
		static_cast<const uint32_t*>(ctx->d_aux[1]), static_cast<const uint32_t*>(ctx->d_aux[2]), static_cast<const uint32_t*>(ctx->d_aux[3]), n_big, total_clusters,
		static_cast<uint64_t*>(ctx->d_aux[0]), perceptual, total_perms, ctx->etc_flavour, (uint32_t)ctx->rank, world);
	return finish(ctx, pOutput_blocks, 0, (size_t)total_clusters * 8, B200_STAT_ETC1S_ENDPOINT_CLUSTERS);
}

extern "C" int b200_etc1s_optimize_selector_codebook(b200_context* ctx, const void* pEtc_blocks, uint32_t total_clusters, const uint32_t* pCluster_offsets,
	const uint32_t* pCluster_block_indices, uint32_t* pOutput_selectors, int perceptual)
{
	ETC_CHECK_BLOCKS(ctx, "b200_etc1s_optimize_selector_codebook");
	if (!total_clusters) return 1;
	if (!pEtc_blocks || !pCluster_offsets || !pCluster_block_indices || !pOutput_selectors) { ctx->fail("b200_etc1s_optimize_selector_codebook: null buffer"); return 0; }
	const uint32_t n = ctx->etc_total_blocks, total_indices = pCluster_offsets[total_clusters];
	for (uint32_t i = 0; i < total_indices; i++)
		if (pCluster_block_indices[i] >= n) { ctx->fail("b200_etc1s_optimize_selector_codebook: block index out of range"); return 0; }
	if (!upload(ctx, 1, pEtc_blocks, (size_t)n * 8)) return 0;
	if (!upload(ctx, 2, pCluster_offsets, ((size_t)total_clusters + 1) * 4)) return 0;
	if (!upload(ctx, 3, pCluster_block_indices, (size_t)total_indices * 4)) return 0;
	if (!ctx->reserve(ctx->d_aux[0], ctx->aux_cap[0], (size_t)total_clusters * 4)) return 0;
	ctx->launches = 1; __atomic_add_fetch(&g_b200_total_launches, 1, __ATOMIC_RELAXED);
	if (!shard_prepare_output(ctx, 0, (size_t)total_clusters * 4)) return 0;
	B200_CUDA_OK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
	const uint32_t warps_per_cta = 4, world = (uint32_t)(ctx->world > 1 ? ctx->world : 1);
	const uint32_t my_clusters = (total_clusters + world - 1) / world;
	k_etc1s_selector_codebook<<<(my_clusters + warps_per_cta - 1) / warps_per_cta, warps_per_cta * 32, 0, ctx->stream>>>(static_cast<const uint32_t*>(ctx->d_etc_blocks),
		static_cast<const uint2*>(ctx->d_aux[1]), static_cast<const uint32_t*>(ctx->d_aux[2]), static_cast<const uint32_t*>(ctx->d_aux[3]), total_clusters,
		static_cast<uint32_t*>(ctx->d_aux[0]), perceptual, (uint32_t)ctx->rank, world);
	return finish(ctx, pOutput_selectors, 0, (size_t)total_clusters * 4, B200_STAT_ETC1S_SELECTOR_CODEBOOK);
}

extern "C" int b200_etc1s_reoptimize_endpoint_clusters(b200_context* ctx, uint32_t total_clusters, const uint32_t* pCluster_offsets, const uint32_t* pCluster_block_indices,
	const uint32_t* pBlock_selectors, const void* pCluster_color5_inten, void* pOut_color5_inten, uint64_t* pOut_new_err, uint64_t* pOut_cur_err, int perceptual, uint32_t total_perms)
{
	ETC_CHECK_BLOCKS(ctx, "b200_etc1s_reoptimize_endpoint_clusters");
	if (!total_clusters) return 1;
	if (total_perms > 165) { ctx->fail("b200_etc1s_reoptimize_endpoint_clusters: total_perms > 165"); return 0; }
	if (!pCluster_offsets || !pCluster_block_indices || !pCluster_color5_inten || !pOut_color5_inten || !pOut_new_err || !pOut_cur_err)
	{ ctx->fail("b200_etc1s_reoptimize_endpoint_clusters: null buffer"); return 0; }
	const uint32_t total_indices = pCluster_offsets[total_clusters];
	for (uint32_t i = 0; i < total_indices; i++)
		if (pCluster_block_indices[i] >= ctx->etc_total_blocks) { ctx->fail("b200_etc1s_reoptimize_endpoint_clusters: block index out of range"); return 0; }
	uint32_t n_big = 0;
	uint32_t* order = cluster_launch_order(pCluster_offsets, total_clusters, n_big);
	if (!order) { ctx->fail("b200_etc1s_reoptimize_endpoint_clusters: out of host memory"); return 0; }
	bool ok = upload(ctx, 1, pCluster_offsets, ((size_t)total_clusters + 1) * 4) && upload(ctx, 2, pCluster_block_indices, (size_t)total_indices * 4) && upload(ctx, 3, order, (size_t)total_clusters * 4) &&
		(!pBlock_selectors || upload(ctx, 4, pBlock_selectors, (size_t)total_indices * 4)) && upload(ctx, 5, pCluster_color5_inten, (size_t)total_clusters * 4);
	if (ok) ok = ctx->reserve(ctx->d_aux[0], ctx->aux_cap[0], (size_t)total_clusters * sizeof(reopt_out));
	if (ok && cudaStreamSynchronize(ctx->stream) != cudaSuccess) ok = false; // `order` is pageable host memory about to be freed
	free(order);
	if (!ok) return 0;
	ctx->launches = 0;
	if (!shard_prepare_output(ctx, 0, (size_t)total_clusters * sizeof(reopt_out))) return 0;
	B200_CUDA_OK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
	const uint32_t warps_per_cta = 4, world = (uint32_t)(ctx->world > 1 ? ctx->world : 1);
	const uint32_t* d_blocks = static_cast<const uint32_t*>(ctx->d_etc_blocks);
	const uint32_t* d_off = static_cast<const uint32_t*>(ctx->d_aux[1]); const uint32_t* d_idx = static_cast<const uint32_t*>(ctx->d_aux[2]);
	const uint32_t* d_order = static_cast<const uint32_t*>(ctx->d_aux[3]); const uint32_t* d_sels = pBlock_selectors ? static_cast<const uint32_t*>(ctx->d_aux[4]) : nullptr;
	const uint32_t* d_cur = static_cast<const uint32_t*>(ctx->d_aux[5]);
	reopt_out* d_out = static_cast<reopt_out*>(ctx->d_aux[0]);
	if (n_big)
	{
		k_etc1s_reoptimize_clusters_big<<<(n_big + world - 1) / world, ETC1S_CTA_TEAM, 0, ctx->stream>>>(d_blocks, d_off, d_idx, d_sels, d_cur, d_order, n_big, d_out, perceptual, total_perms, ctx->etc_flavour, (uint32_t)ctx->rank, world);
		ctx->launches++; __atomic_add_fetch(&g_b200_total_launches, 1, __ATOMIC_RELAXED);
	}
	const uint32_t my_clusters = (total_clusters - n_big + world - 1) / world;
	if (my_clusters)
	{
		k_etc1s_reoptimize_clusters<<<(my_clusters + warps_per_cta - 1) / warps_per_cta, warps_per_cta * 32, 0, ctx->stream>>>(d_blocks, d_off, d_idx, d_sels, d_cur, d_order, n_big, total_clusters, d_out,
			perceptual, total_perms, ctx->etc_flavour, (uint32_t)ctx->rank, world);
		ctx->launches++; __atomic_add_fetch(&g_b200_total_launches, 1, __ATOMIC_RELAXED);
	}
	reopt_out* h = static_cast<reopt_out*>(malloc((size_t)total_clusters * sizeof(reopt_out)));
	if (!h) { ctx->fail("b200_etc1s_reoptimize_endpoint_clusters: out of host memory"); return 0; }
	const int fin = finish(ctx, h, 0, (size_t)total_clusters * sizeof(reopt_out), B200_STAT_ETC1S_REOPTIMIZE_CLUSTERS);
	if (fin)
	{
		uint8_t* c5i = static_cast<uint8_t*>(pOut_color5_inten);
		for (uint32_t c = 0; c < total_clusters; c++)
		{
			const uint32_t w = (uint32_t)h[c].packed; // etc_block bytes 0..3: R5<<3, G5<<3, B5<<3, inten0<<5 | ...
			c5i[c * 4 + 0] = (uint8_t)((w >> 3) & 31); c5i[c * 4 + 1] = (uint8_t)((w >> 11) & 31); c5i[c * 4 + 2] = (uint8_t)((w >> 19) & 31); c5i[c * 4 + 3] = (uint8_t)((w >> 29) & 7);
			pOut_new_err[c] = h[c].new_err; pOut_cur_err[c] = h[c].cur_err;
		}
	}
	free(h);
	return fin;
}

extern "C" int b200_etc1s_subblock_errors(b200_context* ctx, const void* pBlock_color5_inten, uint64_t* pOut_errors, int perceptual)
{
	ETC_CHECK_BLOCKS(ctx, "b200_etc1s_subblock_errors");
	if (!pBlock_color5_inten || !pOut_errors) { ctx->fail("b200_etc1s_subblock_errors: null buffer"); return 0; }
	const uint32_t n = ctx->etc_total_blocks;
	if (!upload(ctx, 1, pBlock_color5_inten, (size_t)n * 4)) return 0;
	if (!ctx->reserve(ctx->d_aux[0], ctx->aux_cap[0], (size_t)n * 8)) return 0;
	ctx->launches = 1; __atomic_add_fetch(&g_b200_total_launches, 1, __ATOMIC_RELAXED);
	if (!shard_prepare_output(ctx, 0, (size_t)n * 8)) return 0;
	B200_CUDA_OK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
	uint32_t first, last;
	shard_range(ctx, n, first, last);
	if (last > first)
		k_etc1s_subblock_errors<<<(last - first + 127) / 128, 128, 0, ctx->stream>>>(static_cast<const uint4*>(ctx->d_etc_blocks), static_cast<const uint32_t*>(ctx->d_aux[1]), first, last,
			static_cast<uint2*>(ctx->d_aux[0]), perceptual);
	uint32_t* h = static_cast<uint32_t*>(malloc((size_t)n * 8));
	if (!h) { ctx->fail("b200_etc1s_subblock_errors: out of host memory"); return 0; }
	const int fin = finish(ctx, h, 0, (size_t)n * 8, B200_STAT_ETC1S_REOPTIMIZE_CLUSTERS);
	if (fin) for (size_t i = 0; i < (size_t)n * 2; i++) pOut_errors[i] = h[i];
	free(h);
	return fin;
}

// ---- ETC1S backend, endpoint prediction (basisu_backend::create_encoder_blocks, encoder/basisu_backend.cpp:405-617) ------------------
// Per block in raster order the reference either predicts the block's endpoint index from its left / upper / upper-left neighbour
// (equal index), or -- endpoint RDO -- replaces it by a neighbour's index if that keeps the block's error within
// endpoint_rdo_quality_thresh times its current error. The neighbours' indices are the ones ALREADY decided, so block (x, y)
// depends on (x-1, y), (x, y-1), (x-1, y-1): blocks on one anti-diagonal x + y = d are independent. One CTA per slice walks the
// diagonals with a barrier between them; the index array lives in global memory and is read back through L2 (__ldcg).
//
// pred_out: 0..2 = predictor, 3 = none (basist::NO_ENDPOINT_PRED_INDEX); bit 7 marks "none, and the block's error was zero" (the one
// case the reference's hit / miss statistics do not count).
struct backend_slice { uint32_t first_block, nbx, nby; };

__global__ void __launch_bounds__(1024) k_etc1s_endpoint_prediction(const uint4* __restrict__ blocks, const uint2* __restrict__ etc_blocks, const backend_slice* __restrict__ slices,
	const uint32_t* __restrict__ c5i, uint32_t* idx, uint8_t* __restrict__ pred_out, float thresh, int perceptual_i)
{
	const backend_slice sl = slices[blockIdx.x];
	const bool perceptual = perceptual_i != 0;
	const bu_tables* T = &d_tables;
	const uint32_t diagonals = sl.nbx + sl.nby - 1;
	for (uint32_t d = 0; d < diagonals; d++)
	{
		const uint32_t y_lo = d >= sl.nbx ? d - (sl.nbx - 1) : 0, y_hi = min(d, sl.nby - 1);
		for (uint32_t y = y_lo + threadIdx.x; y <= y_hi; y += blockDim.x)
		{
			const uint32_t x = d - y, b = sl.first_block + x + y * sl.nbx;
			const uint32_t own = __ldcg(idx + b);
			const bool has[3] = { x > 0, y > 0, x > 0 && y > 0 };
			uint32_t nb[3];
			nb[0] = has[0] ? __ldcg(idx + b - 1) : 0xFFFFFFFFu;
			nb[1] = has[1] ? __ldcg(idx + b - sl.nbx) : 0xFFFFFFFFu;
			nb[2] = has[2] ? __ldcg(idx + b - sl.nbx - 1) : 0xFFFFFFFFu;
			uint32_t pred = etc1s_predict_from_neighbours(own, nb, has);
			if (pred == 3 && thresh > 0.0f)
			{
				uint32_t px[16];
				for (int q = 0; q < 4; q++) { const uint4 v = __ldg(blocks + (size_t)b * 4 + q); px[q * 4] = v.x; px[q * 4 + 1] = v.y; px[q * 4 + 2] = v.z; px[q * 4 + 3] = v.w; }
				const uint2 etc = __ldg(etc_blocks + b);
				uint32_t nb_c5i[3];
				for (uint32_t p = 0; p < 3; p++) nb_c5i[p] = has[p] ? __ldg(c5i + nb[p]) : 0u;
				uint32_t new_index = own;
				pred = etc1s_endpoint_rdo(T, perceptual, px, etc.x, etc.y, nb, nb_c5i, has, thresh, new_index);
				if ((pred & 3u) != 3u) __stcg(idx + b, new_index);
			}
			pred_out[b] = (uint8_t)pred;
		}
		__syncthreads();
	}
}

extern "C" int b200_etc1s_backend_endpoint_prediction(b200_context* ctx, uint32_t num_slices, const uint32_t* pSlice_first_block_nbx_nby, const void* pEtc_blocks,
	uint32_t total_endpoints, const void* pEndpoint_color5_inten, float endpoint_rdo_quality_thresh, int perceptual, uint32_t* pBlock_endpoint_indices, uint8_t* pOut_predictors)
{
	ETC_CHECK_BLOCKS(ctx, "b200_etc1s_backend_endpoint_prediction");
	if (!num_slices || !pSlice_first_block_nbx_nby || !pEtc_blocks || !pEndpoint_color5_inten || !pBlock_endpoint_indices || !pOut_predictors || !total_endpoints)
	{ ctx->fail("b200_etc1s_backend_endpoint_prediction: null or empty input"); return 0; }
	const uint32_t n = ctx->etc_total_blocks;
	for (uint32_t s = 0; s < num_slices; s++)
	{
		const uint32_t first = pSlice_first_block_nbx_nby[s * 3], nbx = pSlice_first_block_nbx_nby[s * 3 + 1], nby = pSlice_first_block_nbx_nby[s * 3 + 2];
		if (!nbx || !nby || (uint64_t)first + (uint64_t)nbx * nby > n) { ctx->fail("b200_etc1s_backend_endpoint_prediction: slice outside the resident blocks"); return 0; }
	}
	for (uint32_t i = 0; i < n; i++)
		if (pBlock_endpoint_indices[i] >= total_endpoints) { ctx->fail("b200_etc1s_backend_endpoint_prediction: endpoint index out of range"); return 0; }
	if (!upload(ctx, 1, pEtc_blocks, (size_t)n * 8) || !upload(ctx, 2, pSlice_first_block_nbx_nby, (size_t)num_slices * 12) || !upload(ctx, 3, pEndpoint_color5_inten, (size_t)total_endpoints * 4) ||
		!upload(ctx, 4, pBlock_endpoint_indices, (size_t)n * 4)) return 0;
	if (!ctx->reserve(ctx->d_aux[0], ctx->aux_cap[0], (size_t)n)) return 0;
	ctx->launches = 1; __atomic_add_fetch(&g_b200_total_launches, 1, __ATOMIC_RELAXED);
	B200_CUDA_OK(ctx, cudaMemsetAsync(ctx->d_aux[0], 3, n, ctx->stream)); // blocks outside every slice: no predictor
	B200_CUDA_OK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
	k_etc1s_endpoint_prediction<<<num_slices, 1024, 0, ctx->stream>>>(static_cast<const uint4*>(ctx->d_etc_blocks), static_cast<const uint2*>(ctx->d_aux[1]),
		static_cast<const backend_slice*>(ctx->d_aux[2]), static_cast<const uint32_t*>(ctx->d_aux[3]), static_cast<uint32_t*>(ctx->d_aux[4]), static_cast<uint8_t*>(ctx->d_aux[0]),
		endpoint_rdo_quality_thresh, perceptual);
	B200_CUDA_OK(ctx, cudaGetLastError());
	B200_CUDA_OK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
	// every rank of a communicator runs the whole (sequential-by-diagonal) pass: nothing to merge
	B200_CUDA_OK(ctx, cudaMemcpyAsync(pBlock_endpoint_indices, ctx->d_aux[4], (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	B200_CUDA_OK(ctx, cudaMemcpyAsync(pOut_predictors, ctx->d_aux[0], (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
	B200_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
	B200_CUDA_OK(ctx, cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
	ctx->account(B200_STAT_ETC1S_BACKEND_PREDICTION);
	return 1;
}
