// b200_internal.h -- context object behind the C ABI (include/basisu_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/basisu_b200.h"

extern unsigned long long g_b200_total_launches;
void g_b200_stat_add(int stat_id, float ms, uint32_t launches);
struct b200_context;
bool b200_merge_u32(b200_context* ctx, void* d_buf, size_t count); // b200_dist.cu: in-place SUM all-reduce over the ranks (no-op for one rank)
void b200_comm_release(b200_context* ctx);

struct b200_context
{
	int device;
	cudaStream_t stream;
	cudaEvent_t ev0, ev1, ev_t0, ev_t1;
	cudaStream_t copy_in, copy_out; cudaEvent_t pipe_ev[64]; // host-pointer UASTC encode: H2D / D2H streams and per-piece events (created on first use)
	cudaEvent_t stage_ev[64 * 4]; uint32_t stage_ev_count, stage_ev_used; // 4 events per chunk: before k0, k1, k2, after k2
	float stage_ms[3];
	float last_ms;
	uint32_t launches;
	char err[256];

	// grow-only device scratch
	void* d_in; size_t in_cap;       // source blocks (host-pointer entry points)
	void* d_out; size_t out_cap;     // encoded blocks
	void* d_meta; size_t meta_cap;   // per-block class + partition ranks
	void* d_cands; size_t cands_cap; // candidate records, slot-major
	void* d_errs; size_t errs_cap;   // (uastc_err, bc7_err) per candidate, slot-major: the selection's input
	void* d_slots; size_t slots_cap; // slot schedule
	void* d_lists; size_t lists_cap; // per slot-class compacted block lists + their counters
	void* d_aux[6]; size_t aux_cap[6]; // ETC1S stage inputs/outputs, RDO state

	// ETC1S: source blocks of the current slice (b200_etc1s_set_pixel_blocks)
	void* d_etc_blocks; size_t etc_blocks_cap; uint32_t etc_total_blocks;
	int etc_flavour; // B200_ETC1S_FLAVOUR_*
	float stat_ms[B200_STAT_COUNT]; uint32_t stat_launches[B200_STAT_COUNT], stat_calls[B200_STAT_COUNT];
	void* tsvq; // b200_tsvq.cu: host-side result storage + device scratch of b200_tsvq_generate
	void* comm; int rank, world; // b200_dist.cu: NCCL communicator (world <= 1: none)
	float comm_ms; unsigned long long comm_bytes; uint32_t comm_calls;

	void account(int stat_id) // adds the call that just succeeded (last_ms, launches) to the per-family totals
	{
		stat_ms[stat_id] += last_ms; stat_launches[stat_id] += launches; stat_calls[stat_id]++;
		g_b200_stat_add(stat_id, last_ms, launches);
	}
	void fail(const char* msg) { snprintf(err, sizeof(err), "%s", msg); }
	void fail_cuda(const char* what, cudaError_t e) { snprintf(err, sizeof(err), "%s: %s", what, cudaGetErrorString(e)); }
	bool activate()
	{
		cudaError_t e = cudaSetDevice(device);
		if (e != cudaSuccess) { fail_cuda("cudaSetDevice", e); return false; }
		return true;
	}
	cudaEvent_t* chunk_events() // 4 events for the next chunk, or nullptr when the pool is exhausted
	{
		if (stage_ev_used + 4 > 64 * 4) return nullptr;
		while (stage_ev_count < stage_ev_used + 4)
		{
			if (cudaEventCreate(&stage_ev[stage_ev_count]) != cudaSuccess) return nullptr;
			stage_ev_count++;
		}
		cudaEvent_t* e = stage_ev + stage_ev_used;
		stage_ev_used += 4;
		return e;
	}
	void collect_stage_times()
	{
		stage_ms[0] = stage_ms[1] = stage_ms[2] = 0;
		for (uint32_t c = 0; c + 4 <= stage_ev_used; c += 4)
			for (int s = 0; s < 3; s++)
			{
				float ms = 0;
				if (cudaEventElapsedTime(&ms, stage_ev[c + s], stage_ev[c + s + 1]) == cudaSuccess) stage_ms[s] += ms;
			}
	}
	bool reserve(void*& p, size_t& cap, size_t bytes)
	{
		if (bytes <= cap) return true;
		// Stream-ordered allocation from the device's default pool (release threshold raised in b200_create_context): a compressor
		// that creates one context per image (basis_compressor does) gets its buffers back from the pool instead of paying
		// cudaMalloc / cudaFree (device-synchronising, milliseconds each) for every image.
		if (p) { cudaFreeAsync(p, stream); p = nullptr; cap = 0; }
		cudaError_t e = cudaMallocAsync(&p, bytes, stream);
		if (e != cudaSuccess) { p = nullptr; fail_cuda("cudaMallocAsync", e); return false; }
		cap = bytes;
		return true;
	}
};

#define B200_CUDA_OK(ctx, expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { (ctx)->fail_cuda(#expr, e_); return 0; } } while (0)
