// b200_rdo.cu -- UASTC RDO post-pass on sm_100a: b200_uastc_rdo (include/basisu_b200.h), the batch form of
// basisu::uastc_rdo (encoder/basisu_uastc_enc.h:139; implementation uastc_enc.cpp:3824-4164).
//
//   k_rdo_chain    one 256-thread CTA per chain; all chains of all slices of a batch in ONE launch (b200_uastc_rdo_batch: 8 tiles x 4
//                  chains = 32 concurrent CTAs instead of 4) (the reference splits each slice into `total_jobs` contiguous chains,
//                  uastc_enc.cpp:4116-4154, and the result depends on that split, so it is reproduced exactly).
//                  Per block: every thread scores the splice of one of the previous <= 256 blocks' selector bits
//                  (bu_rdo.h::rdo_trial); a block-wide (cost, distance) arg-min picks the winner with the reference's
//                  "first strictly smaller, scanning backwards" rule; the winning thread refits mode-0 endpoints and
//                  writes the block (hint fields neutral). The selector history (std::unordered_map in the reference:
//                  bit pattern -> latest block that registered it) is an exact open-addressing table in global memory,
//                  written by one thread per step and read by all.
//   k_rdo_rehint   one thread per modified block: uastc_recompute_hints (uastc_enc.cpp:3647) == unpack + finish_block.
#include "b200_internal.h"
#include "bu_rdo.h"

using namespace bu;

#include "b200_tables.cuh"

struct hist_entry { uint64_t sel; uint32_t ofs_plus1; uint32_t index; };

__device__ __forceinline__ uint32_t hist_hash(uint32_t ofs, uint64_t sel)
{
	uint64_t h = sel * 0x9E3779B97F4A7C15ull + (uint64_t)ofs * 0xC2B2AE3D27D4EB4Full;
	h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
	return (uint32_t)h;
}
// returns the registered index or -1
__device__ __forceinline__ int hist_find(const hist_entry* __restrict__ tab, uint32_t mask, uint32_t ofs, uint64_t sel)
{
	for (uint32_t i = hist_hash(ofs, sel) & mask;; i = (i + 1) & mask)
	{
		const hist_entry e = tab[i];
		if (!e.ofs_plus1) return -1;
		if (e.ofs_plus1 == ofs + 1 && e.sel == sel) return (int)e.index;
	}
}
__device__ __forceinline__ void hist_set(hist_entry* tab, uint32_t mask, uint32_t ofs, uint64_t sel, uint32_t index)
{
	for (uint32_t i = hist_hash(ofs, sel) & mask;; i = (i + 1) & mask)
	{
		hist_entry e = tab[i];
		if (!e.ofs_plus1 || (e.ofs_plus1 == ofs + 1 && e.sel == sel))
		{
			e.sel = sel; e.ofs_plus1 = ofs + 1; e.index = index;
			tab[i] = e;
			return;
		}
	}
}

__device__ __forceinline__ block_bits ld_bits(const uint4* p) { const uint4 v = *p; block_bits b; b.lo = v.x | ((uint64_t)v.y << 32); b.hi = v.z | ((uint64_t)v.w << 32); return b; }
__device__ __forceinline__ void st_bits(uint4* p, const block_bits& b) { uint4 v; v.x = (uint32_t)b.lo; v.y = (uint32_t)(b.lo >> 32); v.z = (uint32_t)b.hi; v.w = (uint32_t)(b.hi >> 32); *p = v; }

#define RDO_THREADS 256

// chains[c] = { first block, one past the last block } of chain c in the (possibly multi-slice) block array
__global__ void __launch_bounds__(RDO_THREADS) k_rdo_chain(uint4* blocks, const uint4* __restrict__ pixels, const uint2* __restrict__ chains,
	hist_entry* hist_all, uint32_t hist_cap, uint8_t* modified, rdo_params p, int* failed)
{
	const bu_tables* T = &d_tables;
	const uint32_t first = chains[blockIdx.x].x;
	const uint32_t last = chains[blockIdx.x].y;
	hist_entry* hist = hist_all + (size_t)blockIdx.x * hist_cap;
	const uint32_t hmask = hist_cap - 1;
	const uint32_t tid = threadIdx.x;
	const int window = (int)(p.lz_dict_size / 16 > 1 ? p.lz_dict_size / 16 : 1);

	__shared__ float s_t[RDO_THREADS / 32];
	__shared__ uint32_t s_i[RDO_THREADS / 32];
	__shared__ uint32_t s_winner;
	// Everything a step needs that depends on the block alone (its unpacked form, source texels, error, smoothness scale,
	// selector field) is independent of the steps before it: each of the CTA's 8 warps prepares one of the next 8 blocks, then
	// the 8 steps run in order with all 256 threads on the trials. (Measured alternative, round 2: one block per THREAD, 256 per
	// batch -- 4096^2 level 2 with 4 chains 5.98 s instead of 4.69 s: consecutive blocks use different modes, so a warp of 32
	// per-thread prologues serialises ~10 divergent paths, which costs more than 8 warp-uniform prologues per 8 blocks.)
	enum { PRO_OK = 0, PRO_SOLID = 1, PRO_SKIP = 2, PRO_BAD = 3 };
	struct prologue { rdo_step st; uint32_t px[16]; uint32_t state; };
	__shared__ prologue s_pro[RDO_THREADS / 32];

	for (uint32_t base = first; base < last; base += RDO_THREADS / 32)
	{
		{
			const uint32_t bi = base + (tid >> 5);
			if (bi < last)
			{
				rdo_step st;
				uint32_t px[16];
				uint32_t state = PRO_OK;
				st.bits = ld_bits(blocks + bi);
				if (!unpack_block_bits(T, st.bits, st.cur)) state = PRO_BAD;
				else if (st.cur.mode == 8) state = PRO_SOLID;
				else
				{
					const uint4* pp = pixels + (size_t)bi * 4;
#pragma unroll
					for (int r = 0; r < 4; r++) { const uint4 v = __ldg(pp + r); px[r * 4] = v.x; px[r * 4 + 1] = v.y; px[r * 4 + 2] = v.z; px[r * 4 + 3] = v.w; }
					st.smooth_scale = rdo_smooth_scale(p, px);
					bc7_endpoints_of(T, st.cur, st.bc7);
					const uint64_t cur_err = rdo_block_error(T, st.cur, st.bc7, px);
					st.cur_ms_err = (float)cur_err * (1.0f / 64.0f);
					st.cur_rms_err = sqrtf(st.cur_ms_err);
					mode_selector_field(st.cur.mode, st.first_sel_bit, st.total_sel_bits);
					const uint32_t n0 = st.total_sel_bits < 64 ? st.total_sel_bits : 64;
					st.cur_sel_bits = bits_read(st.bits, st.first_sel_bit, n0);
					if (st.cur_rms_err >= p.skip_block_rms_thresh) state = PRO_SKIP;
				}
				if ((tid & 31) == 0)
				{
					prologue& P = s_pro[tid >> 5];
					P.state = state;
					if (state == PRO_OK || state == PRO_SKIP)
					{
						P.st = st;
						for (int k = 0; k < 16; k++) P.px[k] = px[k];
					}
				}
			}
		}
		__syncthreads();

	for (uint32_t bi = base; bi < last && bi < base + RDO_THREADS / 32; bi++)
	{
		const prologue& P = s_pro[bi - base];
		if (P.state == PRO_BAD) { if (!tid) *failed = 1; return; }
		if (P.state == PRO_SOLID) continue;
		// read in place from shared memory (s_pro is only rewritten after the barrier that ends the batch): copying the 270-byte
		// record into every thread's local memory was a quarter of a step's memory traffic
		const rdo_step& st = P.st;
		const uint32_t* px = P.px;
		const uint32_t n0 = st.total_sel_bits < 64 ? st.total_sel_bits : 64;

		if (P.state == PRO_SKIP)
		{
			if (!tid) hist_set(hist, hmask, st.first_sel_bit, st.cur_sel_bits, bi);
			__syncthreads();
			continue;
		}

		const int found = hist_find(hist, hmask, st.first_sel_bit, st.cur_sel_bits);
		const int cur_bits = (found < 0) ? (int)((st.total_sel_bits * p.lz_literal_cost) / 100) : (int)match_cost_estimate((bi - (uint32_t)found) * 16);
		const float base_t = st.cur_ms_err * st.smooth_scale + (float)cur_bits * p.lambda;

		// one trial per thread: prev = bi - 1 - tid (window <= 256 = blockDim; larger dictionaries loop)
		float my_t = 3.0e38f;
		uint32_t my_rank = 0xFFFFFFFFu; // distance - 1: smaller = scanned earlier by the reference
		block_bits my_bits = st.bits;
		const int first_check = ((int)bi - window > (int)first) ? (int)bi - window : (int)first;
		for (int pi = (int)bi - 1 - (int)tid; pi >= first_check; pi -= RDO_THREADS)
		{
			const block_bits prev = ld_bits(blocks + pi);
			const int m = hist_find(hist, hmask, st.first_sel_bit, bits_read(prev, st.first_sel_bit, n0));
			float t; block_bits tb;
			if (!rdo_trial(T, p, st, px, prev, pi, (m < 0) ? pi : m, (int)bi, t, tb)) continue;
			if (t < my_t) { my_t = t; my_rank = (uint32_t)((int)bi - 1 - pi); my_bits = tb; } // ascending distance per thread: strict < keeps the nearest
		}

		// block-wide arg-min of (t, rank)
		float wt = my_t; uint32_t wr = my_rank;
#pragma unroll
		for (int m = 16; m >= 1; m >>= 1)
		{
			const float ot = __shfl_xor_sync(0xffffffffu, wt, m);
			const uint32_t orr = __shfl_xor_sync(0xffffffffu, wr, m);
			if (ot < wt || (ot == wt && orr < wr)) { wt = ot; wr = orr; }
		}
		if ((tid & 31) == 0) { s_t[tid >> 5] = wt; s_i[tid >> 5] = wr; }
		__syncthreads();
		if (tid == 0)
		{
			float bt = s_t[0]; uint32_t br = s_i[0];
			for (int w = 1; w < RDO_THREADS / 32; w++)
				if (s_t[w] < bt || (s_t[w] == bt && s_i[w] < br)) { bt = s_t[w]; br = s_i[w]; }
			s_winner = (br != 0xFFFFFFFFu && bt < base_t) ? br : 0xFFFFFFFFu;
		}
		__syncthreads();
		const uint32_t winner = s_winner;

		if (winner == 0xFFFFFFFFu)
		{
			if (!tid) hist_set(hist, hmask, st.first_sel_bit, st.cur_sel_bits, bi);
		}
		else if (my_rank == winner)
		{
			// exactly one thread owns rank `winner` (ranks are unique: rank = distance - 1, each distance visited by one thread)
			candidate bc;
			if (!unpack_block_bits(T, my_bits, bc)) { *failed = 1; }
			else
			{
				if (p.endpoint_refinement && st.cur.mode == 0) rdo_refine_mode0(T, bc, px);
				const block_bits nb = pack_without_hints(T, bc);
				st_bits(blocks + bi, nb);
				modified[bi] = 1;
				hist_set(hist, hmask, st.first_sel_bit, bits_read(nb, st.first_sel_bit, n0), bi);
			}
		}
		__syncthreads();
	}
		__syncthreads(); // s_pro is rewritten by the next batch
	}
}

__global__ void __launch_bounds__(128) k_rdo_rehint(uint4* blocks, const uint4* __restrict__ pixels, uint32_t n, const uint8_t* __restrict__ modified, level_opts o, int level, uint32_t flags, int* failed)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || !modified[i]) return;
	candidate c;
	if (!unpack_block_bits(&d_tables, ld_bits(blocks + i), c)) { *failed = 1; return; }
	uint32_t px[16];
	const uint4* pp = pixels + (size_t)i * 4;
#pragma unroll
	for (int r = 0; r < 4; r++) { const uint4 v = __ldg(pp + r); px[r * 4] = v.x; px[r * 4 + 1] = v.y; px[r * 4 + 2] = v.z; px[r * 4 + 3] = v.w; }
	uint8_t b[16];
	finish_block(&d_tables, o, level, flags, px, c, b);
	uint4 v;
	v.x = b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24);
	v.y = b[4] | (b[5] << 8) | (b[6] << 16) | ((uint32_t)b[7] << 24);
	v.z = b[8] | (b[9] << 8) | (b[10] << 16) | ((uint32_t)b[11] << 24);
	v.w = b[12] | (b[13] << 8) | (b[14] << 16) | ((uint32_t)b[15] << 24);
	blocks[i] = v;
}

// Shared by every entry point: blocks and source texels already on the device, `slice_blocks[s]` blocks per slice laid end
// to end. Every slice is cut into chains exactly as uastc_rdo cuts it (uastc_enc.cpp:4116-4133, as if a job pool were supplied),
// and ALL chains of ALL slices run in one launch, one CTA each; the deferred hint pass covers the whole array.
static int rdo_run_device(b200_context* ctx, uint4* d_blocks, const uint4* d_pixels, uint32_t num_slices, const uint32_t* slice_blocks,
	const b200_uastc_rdo_params* params, uint32_t flags, uint32_t total_jobs, int* h_failed_out)
{
	if (!(params->lambda > 0.0f) || !params->lz_dict_size) { ctx->fail("b200_uastc_rdo: lambda and lz_dict_size must be positive"); return 0; }
	rdo_params p;
	p.lz_dict_size = params->lz_dict_size; p.lambda = params->lambda; p.max_allowed_rms_increase_ratio = params->max_allowed_rms_increase_ratio;
	p.skip_block_rms_thresh = params->skip_block_rms_thresh; p.endpoint_refinement = params->endpoint_refinement;
	p.max_smooth_block_std_dev = params->max_smooth_block_std_dev; p.smooth_block_max_error_scale = params->smooth_block_max_error_scale;
	p.lz_literal_cost = params->lz_literal_cost;

	uint64_t total = 0;
	uint32_t max_chain = 0, num_chains = 0;
	for (uint32_t s = 0; s < num_slices; s++)
	{
		const uint32_t n = slice_blocks[s];
		if (!n) continue;
		uint32_t per_job = total_jobs ? n / total_jobs : 0;
		if (total_jobs <= 1 || per_job <= 8) per_job = n;
		num_chains += (n + per_job - 1) / per_job;
		if (per_job > max_chain) max_chain = per_job;
		total += n;
	}
	if (!total) return 1;
	if (total > 0xFFFFFFFFull) { ctx->fail("b200_uastc_rdo: more than 2^32 blocks in one call"); return 0; }
	const uint32_t num_blocks = (uint32_t)total;
	uint2* h_chains = static_cast<uint2*>(malloc((size_t)num_chains * sizeof(uint2)));
	if (!h_chains) { ctx->fail("b200_uastc_rdo: out of host memory"); return 0; }
	{
		uint32_t c = 0, base = 0;
		for (uint32_t s = 0; s < num_slices; s++)
		{
			const uint32_t n = slice_blocks[s];
			if (!n) continue;
			uint32_t per_job = total_jobs ? n / total_jobs : 0;
			if (total_jobs <= 1 || per_job <= 8) per_job = n;
			for (uint32_t f = 0; f < n; f += per_job) { h_chains[c].x = base + f; h_chains[c].y = base + ((f + per_job < n) ? f + per_job : n); c++; }
			base += n;
		}
	}
	uint32_t cap = 64;
	while (cap < 2 * max_chain) cap <<= 1;

	const size_t hist_bytes = (size_t)num_chains * cap * sizeof(hist_entry);
	const size_t flag_ofs = ((size_t)num_blocks + 15) & ~(size_t)15; // per-block "modified" bytes, then the failure flag, then the chain table
	const size_t chains_ofs = flag_ofs + 16;
	bool ok = ctx->reserve(ctx->d_aux[4], ctx->aux_cap[4], hist_bytes) && ctx->reserve(ctx->d_aux[5], ctx->aux_cap[5], chains_ofs + (size_t)num_chains * sizeof(uint2));
	if (!ok) { free(h_chains); return 0; }
	uint8_t* modified = static_cast<uint8_t*>(ctx->d_aux[5]);
	int* failed = reinterpret_cast<int*>(modified + flag_ofs);
	uint2* d_chains = reinterpret_cast<uint2*>(modified + chains_ofs);
	cudaError_t e = cudaMemsetAsync(ctx->d_aux[4], 0, hist_bytes, ctx->stream);
	if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_aux[5], 0, chains_ofs, ctx->stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(d_chains, h_chains, (size_t)num_chains * sizeof(uint2), cudaMemcpyHostToDevice, ctx->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream); // h_chains is pageable and freed next
	free(h_chains);
	if (e != cudaSuccess) { ctx->fail_cuda("b200_uastc_rdo: setup", e); return 0; }
	B200_CUDA_OK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
	k_rdo_chain<<<num_chains, RDO_THREADS, 0, ctx->stream>>>(d_blocks, d_pixels, d_chains, static_cast<hist_entry*>(ctx->d_aux[4]), cap, modified, p, failed);
	const int lvl = (int)(flags & 0xF);
	k_rdo_rehint<<<(num_blocks + 127) / 128, 128, 0, ctx->stream>>>(d_blocks, d_pixels, num_blocks, modified, make_level_opts(lvl), lvl, flags, failed);
	ctx->launches = 2; __atomic_add_fetch(&g_b200_total_launches, 2, __ATOMIC_RELAXED);
	B200_CUDA_OK(ctx, cudaGetLastError());
	B200_CUDA_OK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
	B200_CUDA_OK(ctx, cudaMemcpyAsync(h_failed_out, failed, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	B200_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
	B200_CUDA_OK(ctx, cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
	if (*h_failed_out) { ctx->fail("b200_uastc_rdo: a block failed to unpack (invalid UASTC input)"); return 0; } // reference: cECFailedUASTCRDOPostProcess
	ctx->account(B200_STAT_UASTC_RDO);
	return 1;
}

extern "C" int b200_uastc_rdo_batch_device(b200_context* ctx, uint32_t num_slices, const uint32_t* pSlice_num_blocks, void* dBlocks, const void* dBlock_pixels,
	const b200_uastc_rdo_params* params, uint32_t flags, uint32_t total_jobs)
{
	if (!ctx) return 0;
	if (!ctx->activate()) return 0;
	if (!num_slices) return 1;
	if (!pSlice_num_blocks || !dBlocks || !dBlock_pixels || !params) { ctx->fail("b200_uastc_rdo_batch_device: null argument"); return 0; }
	int h_failed = 0;
	return rdo_run_device(ctx, static_cast<uint4*>(dBlocks), static_cast<const uint4*>(dBlock_pixels), num_slices, pSlice_num_blocks, params, flags, total_jobs, &h_failed);
}

extern "C" int b200_uastc_rdo_batch(b200_context* ctx, uint32_t num_slices, const uint32_t* pSlice_num_blocks, void* pBlocks, const void* pBlock_pixels,
	const b200_uastc_rdo_params* params, uint32_t flags, uint32_t total_jobs)
{
	if (!ctx) return 0;
	if (!ctx->activate()) return 0;
	if (!num_slices) return 1;
	if (!pSlice_num_blocks || !pBlocks || !pBlock_pixels || !params) { ctx->fail("b200_uastc_rdo: null argument"); return 0; }
	uint64_t total = 0;
	for (uint32_t s = 0; s < num_slices; s++) total += pSlice_num_blocks[s];
	if (!total) return 1;
	if (!ctx->reserve(ctx->d_out, ctx->out_cap, (size_t)total * 16)) return 0;
	if (!ctx->reserve(ctx->d_in, ctx->in_cap, (size_t)total * 64)) return 0;
	B200_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_out, pBlocks, (size_t)total * 16, cudaMemcpyHostToDevice, ctx->stream));
	B200_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_in, pBlock_pixels, (size_t)total * 64, cudaMemcpyHostToDevice, ctx->stream));
	int h_failed = 0;
	if (!rdo_run_device(ctx, static_cast<uint4*>(ctx->d_out), static_cast<const uint4*>(ctx->d_in), num_slices, pSlice_num_blocks, params, flags, total_jobs, &h_failed))
		return 0; // pBlocks is left untouched on failure, so a caller may still run its own CPU pass on the encoder's output
	B200_CUDA_OK(ctx, cudaMemcpy(pBlocks, ctx->d_out, (size_t)total * 16, cudaMemcpyDeviceToHost));
	return 1;
}

// The compressor's whole per-slice UASTC path (comp.cpp:1996-2089) for a list of slices in one call: encode, then (params != NULL)
// the RDO post-pass, with the blocks staying in HBM between the two stages. Source blocks in, final UASTC blocks out (HOST pointers).
extern "C" int b200_uastc_encode_rdo_blocks(b200_context* ctx, uint32_t num_slices, const uint32_t* pSlice_num_blocks, const void* pBlocks, void* pOut,
	uint32_t flags, const b200_uastc_rdo_params* params, uint32_t total_jobs)
{
	if (!ctx) return 0;
	if (!ctx->activate()) return 0;
	if (!num_slices) return 1;
	if (!pSlice_num_blocks || !pBlocks || !pOut) { ctx->fail("b200_uastc_encode_rdo_blocks: null argument"); return 0; }
	uint64_t total = 0;
	for (uint32_t s = 0; s < num_slices; s++) total += pSlice_num_blocks[s];
	if (!total) return 1;
	if (total > 0xFFFFFFFFull) { ctx->fail("b200_uastc_encode_rdo_blocks: more than 2^32 blocks in one call"); return 0; }
	if (!ctx->reserve(ctx->d_in, ctx->in_cap, (size_t)total * 64)) return 0;
	if (!ctx->reserve(ctx->d_out, ctx->out_cap, (size_t)total * 16)) return 0;
	B200_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_in, pBlocks, (size_t)total * 64, cudaMemcpyHostToDevice, ctx->stream));
	if (!b200_uastc_encode_blocks_device(ctx, ctx->d_in, (uint32_t)total, ctx->d_out, flags)) return 0;
	const float enc_ms = ctx->last_ms; const uint32_t enc_launches = ctx->launches;
	if (params)
	{
		int h_failed = 0;
		if (!rdo_run_device(ctx, static_cast<uint4*>(ctx->d_out), static_cast<const uint4*>(ctx->d_in), num_slices, pSlice_num_blocks, params, flags, total_jobs, &h_failed)) return 0;
		ctx->last_ms += enc_ms; ctx->launches += enc_launches;
	}
	B200_CUDA_OK(ctx, cudaMemcpy(pOut, ctx->d_out, (size_t)total * 16, cudaMemcpyDeviceToHost));
	return 1;
}

extern "C" int b200_uastc_rdo(b200_context* ctx, uint32_t num_blocks, void* pBlocks, const void* pBlock_pixels,
	const b200_uastc_rdo_params* params, uint32_t flags, uint32_t total_jobs)
{
	return b200_uastc_rdo_batch(ctx, 1, &num_blocks, pBlocks, pBlock_pixels, params, flags, total_jobs);
}
