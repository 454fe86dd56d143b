// b200_internal.h -- context object behind the C ABI (include/basisu_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/basisu_b200.h"

struct b200_context
{
	int device;
	cudaStream_t stream;
	cudaEvent_t ev0, ev1;
	float last_ms;
	uint32_t launches;
	char err[256];

	// grow-only device scratch
	void* d_in; size_t in_cap;       // source blocks (host-pointer entry points)
	void* d_out; size_t out_cap;     // encoded blocks
	void* d_meta; size_t meta_cap;   // per-block class + partition ranks
	void* d_cands; size_t cands_cap; // candidate records, slot-major
	void* d_slots; size_t slots_cap; // slot schedule
	void* d_aux[6]; size_t aux_cap[6]; // ETC1S stage inputs/outputs, RDO state

	// ETC1S: source blocks of the current slice (b200_etc1s_set_pixel_blocks)
	void* d_etc_blocks; size_t etc_blocks_cap; uint32_t etc_total_blocks;

	void fail(const char* msg) { snprintf(err, sizeof(err), "%s", msg); }
	void fail_cuda(const char* what, cudaError_t e) { snprintf(err, sizeof(err), "%s: %s", what, cudaGetErrorString(e)); }
	bool activate()
	{
		cudaError_t e = cudaSetDevice(device);
		if (e != cudaSuccess) { fail_cuda("cudaSetDevice", e); return false; }
		return true;
	}
	bool reserve(void*& p, size_t& cap, size_t bytes)
	{
		if (bytes <= cap) return true;
		if (p) { cudaFree(p); p = nullptr; cap = 0; }
		cudaError_t e = cudaMalloc(&p, bytes);
		if (e != cudaSuccess) { p = nullptr; fail_cuda("cudaMalloc", e); return false; }
		cap = bytes;
		return true;
	}
};

#define B200_CUDA_OK(ctx, expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { (ctx)->fail_cuda(#expr, e_); return 0; } } while (0)
