// b200_uastc.cu -- UASTC LDR 4x4 block encoder kernels for sm_100a and their C-ABI entry points.
//
// Pipeline, all buffers resident in HBM:
//   k_classify_rank   1 thread / block   : 128-bit loads of the 64 B block, class bits, solid-colour fast path (packs the
//                                          block directly), partition ranking for the estimated-partition modes
//   k_candidates      1 thread / (block, candidate slot) : endpoint/selector fit of one (mode, variant), UASTC decode error,
//                                          BC7-transcode error -> 64 B candidate record
//   k_finish          1 thread / block   : windowed arg-min over the block's candidates, BC1/ETC2-EAC/ETC1 hints, bit packing
// Build: nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false (no FMA contraction: the reference is built for baseline
// x86-64, and bit-exactness includes its float/double rounding; SURVEY.md section 7 hard part 1).
#include "b200_internal.h"
#define BU_UASTC_ENCODER_TU 1 // the kernels of this translation unit stage the colour-cell tables (bu_ccc.h)
#include "bu_slots.h"

using namespace bu;

#include "b200_tables.cuh"

// Occupancy knobs (tools/build_variant.sh -D...): minimum resident CTAs per SM the register allocator must allow.
#ifndef BU_CAND_MINB
#define BU_CAND_MINB 10
#endif
#ifndef BU_FIN_MINB
#define BU_FIN_MINB 10
#endif
#ifndef BU_CLS_MINB
#define BU_CLS_MINB 8
#endif

struct block_meta { block_class k; block_ranks ranks; }; // 16 B

__device__ __forceinline__ void load_block(const uint4* __restrict__ blocks, uint32_t i, uint32_t* px)
{
	const uint4* p = blocks + (size_t)i * 4;
#pragma unroll
	for (int r = 0; r < 4; r++)
	{
		const uint4 v = __ldg(p + r);
		px[r * 4 + 0] = v.x; px[r * 4 + 1] = v.y; px[r * 4 + 2] = v.z; px[r * 4 + 3] = v.w;
	}
}

// Work lists: one compacted list of block indices per slot class (LA / RGB / alpha slots), so that every warp of
// k_candidates is fully populated with blocks that actually run its slot. Layout: counts[4] then lists[3][n].
struct work_lists { uint32_t* counts; uint32_t* list[3]; };

__device__ __forceinline__ void list_append(uint32_t* count, uint32_t* list, bool pred, uint32_t value)
{
	// warp-aggregated: one atomic per warp; a warp's blocks stay contiguous and in order
	const uint32_t mask = __ballot_sync(0xFFFFFFFFu, pred);
	if (!mask) return;
	const uint32_t lane = threadIdx.x & 31u, leader = (uint32_t)__ffs((int)mask) - 1u;
	uint32_t base = 0;
	if (lane == leader) base = atomicAdd(count, (uint32_t)__popc(mask));
	base = __shfl_sync(0xFFFFFFFFu, base, (int)leader);
	if (pred) list[base + (uint32_t)__popc(mask & ((1u << lane) - 1u))] = value;
}

__global__ void __launch_bounds__(128, BU_CLS_MINB) k_classify_rank(const uint4* __restrict__ blocks, uint32_t n, block_meta* __restrict__ meta, uint4* __restrict__ out,
	work_lists wl, level_opts o)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	const bool in_range = i < n;
	bool active = false;
	block_meta m;
	m.k.has_alpha = 0; m.k.is_la = 0; m.k.solid = 0;
	if (in_range)
	{
		uint32_t px[16];
		load_block(blocks, i, px);
		m.k = classify_block(px, o.la_only_transparent != 0);
		if (m.k.solid)
		{
			uint8_t b[16];
			pack_solid_block(&d_tables, px[0], b);
			uint4 v;
			v.x = b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24);
			v.y = b[4] | (b[5] << 8) | (b[6] << 16) | ((uint32_t)b[7] << 24);
			v.z = b[8] | (b[9] << 8) | (b[10] << 16) | ((uint32_t)b[11] << 24);
			v.w = b[12] | (b[13] << 8) | (b[14] << 16) | ((uint32_t)b[15] << 24);
			out[i] = v;
			for (int j = 0; j < 12; j++) m.ranks.r[j] = 0;
		}
		else
		{
			rank_block(&d_tables, o, m.k, px, m.ranks);
			active = true;
		}
		meta[i] = m;
	}
	slot_desc probe; probe.mode = 0; probe.variant = 0; probe.pad = 0;
	for (uint32_t c = 0; c < 3; c++)
	{
		probe.klass = (uint8_t)c;
		list_append(wl.counts + c, wl.list[c], active && slot_active(probe, m.k, o), i);
	}
}

// gridDim.y = slot index within the launch's slot class; candidates are stored slot-major (cands[slot][block], 64 B records).
__global__ void __launch_bounds__(128, BU_CAND_MINB) k_candidates(const uint4* __restrict__ blocks, uint32_t n, const block_meta* __restrict__ meta,
	candidate* __restrict__ cands, uint2* __restrict__ errs, const slot_desc* __restrict__ slots, uint32_t first_slot, const uint32_t* __restrict__ list, const uint32_t* __restrict__ count, level_opts o)
{
	const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t si = first_slot + blockIdx.y;
	const slot_desc s = slots[si];
#if defined(BU_STAGE_TABLES) && defined(__CUDA_ARCH__)
	if (blockIdx.x * blockDim.x >= __ldg(count)) return; // whole CTA beyond the work list
	{
		// TMA bulk copies of the mode's four table rows into shared memory, completion on one mbarrier
		__shared__ __align__(8) unsigned long long s_mbar;
		const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
		if (threadIdx.x == 0)
		{
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar));
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		__syncthreads();
		if (threadIdx.x == 0)
		{
			const bu_tables* T = &d_tables;
			const uint32_t wbits = T->mode_weight_bits[s.mode], rs = (uint32_t)T->range_slot[T->mode_endpoint_range[s.mode]];
			const uint32_t dst[4] = { (uint32_t)__cvta_generic_to_shared(bu_stage::s_rows.su), (uint32_t)__cvta_generic_to_shared(bu_stage::s_rows.nearest),
				(uint32_t)__cvta_generic_to_shared(bu_stage::s_rows.wt), (uint32_t)__cvta_generic_to_shared(bu_stage::s_rows.wx) };
			const void* src[4] = { T->sorted_unq + rs * 256, T->nearest + rs * 256, T->weights + wbits * 32, T->weightsx + wbits * 32 * 4 };
			const uint32_t bytes[4] = { 256, 256, 32, 512 };
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(256u + 256u + 32u + 512u) : "memory");
			for (int k = 0; k < 4; k++)
				asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst[k]), "l"(src[k]), "r"(bytes[k]), "r"(mbar) : "memory");
		}
		uint32_t done;
		do
		{
			asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(mbar), "r"(0u) : "memory");
		} while (!done);
	}
#endif
	if (li >= __ldg(count)) return;
	const uint32_t i = __ldg(list + li);
	const block_meta m = meta[i];
	uint32_t px[16];
	load_block(blocks, i, px);
	candidate c;
	run_slot(&d_tables, o, s, m.k, m.ranks, px, c);
	const uint4* src = reinterpret_cast<const uint4*>(&c);
	uint4* d4 = reinterpret_cast<uint4*>(cands + (size_t)si * n + i);
	d4[0] = src[0]; d4[1] = src[1]; d4[2] = src[2]; d4[3] = src[3];
	errs[(size_t)si * n + i] = make_uint2(c.uastc_err, c.bc7_err); // what the selection in k_finish reads: 8 B per slot, coalesced across blocks
}

// One thread per block. (A cooperative 8-lanes-per-block variant was measured in round 1: 33.5 ms vs 23.8 ms for this
// one -- the stage is issue-bound, not latency-bound, so spreading a block over lanes only adds redundant instructions.)
// A slot's record exists iff the slot is active for the block's class (same predicate that built the work lists).
// CAP bounds the per-thread selection arrays (thread-local memory): 32 covers levels 0-3, MAX_SLOTS level 4.
template<int CAP> __global__ void __launch_bounds__(128, BU_FIN_MINB) k_finish(const uint4* __restrict__ blocks, uint32_t n, const block_meta* __restrict__ meta,
	const candidate* __restrict__ cands, const uint2* __restrict__ errs, const slot_desc* __restrict__ slots, uint32_t nslots, uint4* __restrict__ out, level_opts o, int level, uint32_t flags)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const block_meta m = meta[i];
	if (m.k.solid) return;

	uint32_t ue[CAP], be[CAP];
	uint8_t modes[CAP], idx[CAP];
	uint32_t cnt = 0;
	for (uint32_t s = 0; s < nslots; s++)
	{
		const slot_desc sd = slots[s];
		if (!slot_active(sd, m.k, o)) continue;
		const uint2 e = __ldg(errs + (size_t)s * n + i);
		ue[cnt] = e.x; be[cnt] = e.y; modes[cnt] = sd.mode; idx[cnt] = (uint8_t)s;
		cnt++;
	}
	const int best = select_candidate(cnt, ue, be, modes, flags);

	candidate c;
	{
		const uint4* src = reinterpret_cast<const uint4*>(cands + (size_t)idx[best] * n + i);
		uint4* d4 = reinterpret_cast<uint4*>(&c);
		d4[0] = src[0]; d4[1] = src[1]; d4[2] = src[2]; d4[3] = src[3];
	}
	uint32_t px[16];
	load_block(blocks, i, px);
	uint8_t b[16];
	finish_block(&d_tables, o, level, flags, px, c, b);
	uint4 v;
	v.x = b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24);
	v.y = b[4] | (b[5] << 8) | (b[6] << 16) | ((uint32_t)b[7] << 24);
	v.z = b[8] | (b[9] << 8) | (b[10] << 16) | ((uint32_t)b[11] << 24);
	v.w = b[12] | (b[13] << 8) | (b[14] << 16) | ((uint32_t)b[15] << 24);
	out[i] = v;
}

// ---- host side -----------------------------------------------------------------------------------------------------------

static bool uastc_encode_chunk(b200_context* ctx, const uint4* dBlocks, uint32_t n, uint4* dOut, uint32_t flags)
{
	const int level = clampi((int)(flags & 7), 0, 4);
	const level_opts o = make_level_opts(level);
	slot_desc slots[MAX_SLOTS];
	const uint32_t nslots = build_slots(o, slots);

	if (!ctx->reserve(ctx->d_meta, ctx->meta_cap, (size_t)n * sizeof(block_meta))) return false;
	if (!ctx->reserve(ctx->d_cands, ctx->cands_cap, (size_t)n * nslots * sizeof(candidate))) return false;
	if (!ctx->reserve(ctx->d_errs, ctx->errs_cap, (size_t)n * nslots * sizeof(uint2))) return false;
	if (!ctx->reserve(ctx->d_slots, ctx->slots_cap, sizeof(slot_desc) * MAX_SLOTS)) return false;
	B200_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_slots, slots, sizeof(slot_desc) * nslots, cudaMemcpyHostToDevice, ctx->stream));

	// Slots are listed class by class (build_slots: LA, then RGB, then alpha), so each class is one contiguous slot range.
	uint32_t class_first[3] = { 0, 0, 0 }, class_count[3] = { 0, 0, 0 };
	for (uint32_t s = 0; s < nslots; s++)
	{
		const uint32_t k = slots[s].klass;
		if (!class_count[k]) class_first[k] = s;
		class_count[k]++;
	}

	if (!ctx->reserve(ctx->d_lists, ctx->lists_cap, sizeof(uint32_t) * (4 + 3 * (size_t)n))) return false;
	work_lists wl;
	wl.counts = static_cast<uint32_t*>(ctx->d_lists);
	for (int k = 0; k < 3; k++) wl.list[k] = wl.counts + 4 + (size_t)k * n;
	B200_CUDA_OK(ctx, cudaMemsetAsync(wl.counts, 0, sizeof(uint32_t) * 4, ctx->stream));

	const uint32_t tpb = 128, gx = (n + tpb - 1) / tpb;
	block_meta* meta = static_cast<block_meta*>(ctx->d_meta);
	candidate* cands = static_cast<candidate*>(ctx->d_cands);
	uint2* errs = static_cast<uint2*>(ctx->d_errs);
	const slot_desc* d_slots = static_cast<const slot_desc*>(ctx->d_slots);

	cudaEvent_t* ev = ctx->chunk_events();
	if (ev) cudaEventRecord(ev[0], ctx->stream);
	k_classify_rank<<<gx, tpb, 0, ctx->stream>>>(dBlocks, n, meta, dOut, wl, o);
	ctx->launches++; __atomic_add_fetch(&g_b200_total_launches, 1, __ATOMIC_RELAXED);
	if (ev) cudaEventRecord(ev[1], ctx->stream);
	// One launch per slot class over that class's work list. The list length is only known on the device: the grid covers
	// the worst case (every block in the class) and surplus CTAs retire on their first instruction.
	for (int k = 0; k < 3; k++)
	{
		if (!class_count[k]) continue;
		k_candidates<<<dim3(gx, class_count[k]), tpb, 0, ctx->stream>>>(dBlocks, n, meta, cands, errs, d_slots, class_first[k], wl.list[k], wl.counts + k, o);
		ctx->launches++; __atomic_add_fetch(&g_b200_total_launches, 1, __ATOMIC_RELAXED);
	}
	if (ev) cudaEventRecord(ev[2], ctx->stream);
	if (nslots <= 32) k_finish<32><<<gx, tpb, 0, ctx->stream>>>(dBlocks, n, meta, cands, errs, d_slots, nslots, dOut, o, level, flags);
	else k_finish<MAX_SLOTS><<<gx, tpb, 0, ctx->stream>>>(dBlocks, n, meta, cands, errs, d_slots, nslots, dOut, o, level, flags);
	ctx->launches++; __atomic_add_fetch(&g_b200_total_launches, 1, __ATOMIC_RELAXED);
	if (ev) cudaEventRecord(ev[3], ctx->stream);
	B200_CUDA_OK(ctx, cudaGetLastError());
	return true;
}

// Blocks per pass: bounds the candidate scratch (slots * 64 B per block) while keeping every launch many waves deep.
static uint32_t uastc_chunk_blocks(uint32_t flags)
{
	const int level = clampi((int)(flags & 7), 0, 4);
	return (level == 4) ? (1u << 18) : (1u << 20);
}

extern "C" int b200_uastc_encode_blocks_device(b200_context* ctx, const void* dBlocks, uint32_t num_blocks, void* dOut, uint32_t flags)
{
	if (!ctx) return 0;
	if (!ctx->activate()) return 0;
	ctx->launches = 0;
	ctx->stage_ev_used = 0;
	if (!num_blocks) { ctx->last_ms = 0; return 1; }
	B200_CUDA_OK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
	const uint32_t chunk = uastc_chunk_blocks(flags);
	for (uint32_t first = 0; first < num_blocks; first += chunk)
	{
		const uint32_t n = (num_blocks - first < chunk) ? (num_blocks - first) : chunk;
		if (!uastc_encode_chunk(ctx, static_cast<const uint4*>(dBlocks) + (size_t)first * 4, n, static_cast<uint4*>(dOut) + first, flags)) return 0;
	}
	B200_CUDA_OK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
	B200_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
	B200_CUDA_OK(ctx, cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
	ctx->collect_stage_times();
	ctx->account(B200_STAT_UASTC_ENCODE);
	return 1;
}

// Host-pointer form. Large inputs are cut into pieces that flow through three streams: H2D of piece i+1 and D2H of piece i-1 overlap
// the kernels of piece i (two copy engines + the SMs), so the call costs ~ the kernels plus one piece's copy each way instead of
// copy + kernels + copy in series (at level 0 the 80 MiB of PCIe traffic is as long as the kernels). Needs pinned host memory to
// overlap; with pageable memory the copies serialise and the result is the same.
extern "C" int b200_uastc_encode_blocks(b200_context* ctx, const void* pBlocks, uint32_t num_blocks, void* pOut, uint32_t flags)
{
	if (!ctx) return 0;
	if (!ctx->activate()) return 0;
	ctx->launches = 0;
	ctx->stage_ev_used = 0;
	if (!num_blocks) { ctx->last_ms = 0; return 1; }
	if (!pBlocks || !pOut) { ctx->fail("b200_uastc_encode_blocks: null buffer"); return 0; }
	if (!ctx->reserve(ctx->d_in, ctx->in_cap, (size_t)num_blocks * 64)) return 0;
	if (!ctx->reserve(ctx->d_out, ctx->out_cap, (size_t)num_blocks * 16)) return 0;

	const uint32_t chunk = uastc_chunk_blocks(flags);
	// Only where the copies are comparable to the kernels (levels 0 and 1: 4 and 14 ms of kernels per 2^20 blocks against ~2 ms of
	// PCIe): every piece is five launches with their own tail waves, and at level 2 (32 ms of kernels) two pieces gave 34.1 ms
	// against 33.9 ms unpipelined and eight pieces 39.7 ms (measured), so levels >= 2 keep the single copy / kernels / copy sequence.
	uint32_t piece = chunk;
	const int level_ = clampi((int)(flags & 7), 0, 4);
	if (level_ <= 1 && num_blocks >= (1u << 19)) piece = 1u << 18;
	const uint32_t npieces = (num_blocks + piece - 1) / piece;
	const bool pipelined = level_ <= 1 && npieces > 1 && npieces <= 32;
	if (pipelined && !ctx->copy_in)
	{
		if (cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking) != cudaSuccess)
		{ ctx->fail("b200_uastc_encode_blocks: stream creation failed"); return 0; }
		for (int i = 0; i < 64; i++)
			if (cudaEventCreateWithFlags(&ctx->pipe_ev[i], cudaEventDisableTiming) != cudaSuccess) { ctx->fail("b200_uastc_encode_blocks: event creation failed"); return 0; }
	}

	float total_ms = 0;
	B200_CUDA_OK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
	if (!pipelined)
	{
		B200_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_in, pBlocks, (size_t)num_blocks * 64, cudaMemcpyHostToDevice, ctx->stream));
		for (uint32_t first = 0; first < num_blocks; first += chunk)
		{
			const uint32_t n = (num_blocks - first < chunk) ? (num_blocks - first) : chunk;
			if (!uastc_encode_chunk(ctx, static_cast<const uint4*>(ctx->d_in) + (size_t)first * 4, n, static_cast<uint4*>(ctx->d_out) + first, flags)) return 0;
		}
		B200_CUDA_OK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
		B200_CUDA_OK(ctx, cudaMemcpyAsync(pOut, ctx->d_out, (size_t)num_blocks * 16, cudaMemcpyDeviceToHost, ctx->stream));
		B200_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
	}
	else
	{
		// the copy streams start after whatever the main stream had queued (ev0), the main stream ends after the last D2H
		B200_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->copy_in, ctx->ev0, 0));
		for (uint32_t k = 0; k < npieces; k++)
		{
			const uint32_t first = k * piece, n = (num_blocks - first < piece) ? (num_blocks - first) : piece;
			B200_CUDA_OK(ctx, cudaMemcpyAsync(static_cast<uint8_t*>(ctx->d_in) + (size_t)first * 64, static_cast<const uint8_t*>(pBlocks) + (size_t)first * 64, (size_t)n * 64, cudaMemcpyHostToDevice, ctx->copy_in));
			B200_CUDA_OK(ctx, cudaEventRecord(ctx->pipe_ev[k], ctx->copy_in));
		}
		for (uint32_t k = 0; k < npieces; k++)
		{
			const uint32_t first = k * piece, n = (num_blocks - first < piece) ? (num_blocks - first) : piece;
			B200_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->pipe_ev[k], 0));
			if (!uastc_encode_chunk(ctx, static_cast<const uint4*>(ctx->d_in) + (size_t)first * 4, n, static_cast<uint4*>(ctx->d_out) + first, flags)) return 0;
			B200_CUDA_OK(ctx, cudaEventRecord(ctx->pipe_ev[32 + k], ctx->stream));
			B200_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->copy_out, ctx->pipe_ev[32 + k], 0));
			B200_CUDA_OK(ctx, cudaMemcpyAsync(static_cast<uint8_t*>(pOut) + (size_t)first * 16, static_cast<uint8_t*>(ctx->d_out) + (size_t)first * 16, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->copy_out));
		}
		B200_CUDA_OK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
		B200_CUDA_OK(ctx, cudaStreamSynchronize(ctx->copy_out));
		B200_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
	}
	B200_CUDA_OK(ctx, cudaEventElapsedTime(&total_ms, ctx->ev0, ctx->ev1));
	ctx->last_ms = total_ms;
	ctx->collect_stage_times();
	ctx->account(B200_STAT_UASTC_ENCODE);
	return 1;
}
