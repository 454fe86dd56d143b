// b200_backend.cu -- ETC1S backend, palette ordering: basisu::palette_index_reorderer::init (encoder/basisu_enc.cpp:1785-1915) as
// basisu_backend::reoptimize_and_sort_endpoints_codebook calls it (encoder/basisu_backend.cpp:196-198; no distance function).
//
// The reference orders the endpoint palette so that indices which follow each other in the block stream get close numbers (Zeng's
// greedy ordering): hist[a][b] = how often a and b are adjacent in the index stream; start from the most frequent pair; then, N - 2
// times, take the not-yet-placed entry with the largest total count to the placed ones (first one on ties) and put it at the front or
// the back of the placed list, whichever side its counts lean to (a float sum over the placed list, in list order). The reference keeps
// hist as a dense N x N table and rescans it: 0.7 s for 8 000 entries, 4 s for 16 000 on one host core, a large part of the backend.
//
// Here: the adjacency counts are built sparse on the device (radix sort + run-length encode of the adjacent pairs), and one CTA walks
// the N - 2 steps: a block-wide arg-max, the side decision by one thread over the entry's placed neighbours, and a parallel update of
// the neighbours' totals. Every placed neighbour list is kept in placed-list order without sorting: a newly placed entry is always the
// new front or the new back, so it is prepended or appended to each neighbour's list. The side decision adds the same float terms in the
// same order as the reference (terms with a zero count are exact no-ops there), so the ordering is identical.
#include "b200_internal.h"
#include <cub/cub.cuh>
#include <stdlib.h>

namespace {

__global__ void k_pal_pairs(const uint32_t* __restrict__ idx, uint32_t m, unsigned long long* __restrict__ keys)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i + 1 >= m) return;
	const uint32_t a = idx[i], b = idx[i + 1];
	// both directions of every adjacent pair of different indices (inc_hist, enc.h:2816); equal neighbours sort to the end
	keys[2 * (size_t)i] = (a != b) ? (((unsigned long long)a << 32) | b) : ~0ull;
	keys[2 * (size_t)i + 1] = (a != b) ? (((unsigned long long)b << 32) | a) : ~0ull;
}

__global__ void k_pal_rows(const unsigned long long* __restrict__ ukeys, const uint32_t* __restrict__ num_runs, uint32_t* __restrict__ row_ptr, uint32_t n_syms)
{
	// row_ptr[r] = first unique key whose row is >= r (keys ascending; the ~0 run, if any, is excluded)
	const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r > n_syms) return;
	uint32_t nr = *num_runs;
	if (nr && ukeys[nr - 1] == ~0ull) nr--;
	uint32_t lo = 0, hi = nr;
	while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if ((uint32_t)(ukeys[mid] >> 32) < r) lo = mid + 1; else hi = mid; }
	row_ptr[r] = lo;
}

struct pal_state
{
	const unsigned long long* ukeys; const uint32_t* ucount; const uint32_t* row_ptr; // CSR: row r = entries [row_ptr[r], row_ptr[r + 1]), column = low word
	uint32_t* total;      // m_total_count_to_picked
	int* vpos;            // virtual position in the placed list once placed (front inserts count down from -1, back inserts up from 2)
	uint8_t* placed;
	int* nb_vpos; uint32_t* nb_count; uint32_t* n_back; uint32_t* n_front; // per row: placed neighbours in list order (fronts fill the row's tail downwards, backs its head upwards)
	uint32_t* remap;      // out: m_remap_table
	uint32_t n_syms;
	uint32_t* status;     // out: 1 = done, 2 = no adjacent pair of different indices (the caller writes the reference's degenerate result)
};

__device__ __forceinline__ void pal_place(const pal_state& s, uint32_t e, int vp, bool at_back)
{
	// all threads: e joins the placed list at virtual position vp; its unplaced neighbours gain its count and record it in order
	for (uint32_t k = s.row_ptr[e] + threadIdx.x; k < s.row_ptr[e + 1]; k += blockDim.x)
	{
		const uint32_t v = (uint32_t)s.ukeys[k], c = s.ucount[k];
		if (s.placed[v]) continue;
		s.total[v] += c;
		const uint32_t slot = at_back ? (s.row_ptr[v] + s.n_back[v]++) : (s.row_ptr[v + 1] - 1 - s.n_front[v]++);
		s.nb_vpos[slot] = vp; s.nb_count[slot] = c;
	}
	if (!threadIdx.x) { s.placed[e] = 1; s.vpos[e] = vp; }
}

__global__ void __launch_bounds__(1024) k_pal_order(pal_state s, const uint32_t* __restrict__ num_runs)
{
	__shared__ unsigned long long red[32];
	__shared__ uint32_t sh_e; __shared__ int sh_vp; __shared__ int sh_back;
	const uint32_t n = s.n_syms, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t nr = *num_runs;
	if (nr && s.ukeys[nr - 1] == ~0ull) nr--;
	if (!nr) { if (!threadIdx.x) *s.status = 2; return; }

	// find_initial (enc.cpp:1845): the largest count, first in row-major order of (a, b) with a < b
	{
		unsigned long long best = 0; // count << 32 | ~position-in-key-order (keys ascending = row-major order)
		for (uint32_t k = threadIdx.x; k < nr; k += blockDim.x)
		{
			const unsigned long long key = s.ukeys[k];
			if ((uint32_t)(key >> 32) >= (uint32_t)key) continue;
			const unsigned long long cand = ((unsigned long long)s.ucount[k] << 32) | (0xFFFFFFFFu - k);
			best = cand > best ? cand : best;
		}
		for (int o = 16; o; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o); best = t > best ? t : best; }
		if (!lane) red[warp] = best;
		__syncthreads();
		if (!threadIdx.x)
		{
			for (int w = 1; w < 32; w++) best = red[w] > best ? red[w] : best;
			const uint32_t k = 0xFFFFFFFFu - (uint32_t)best;
			sh_e = k;
		}
		__syncthreads();
	}
	const uint32_t k0 = sh_e;
	const uint32_t a = (uint32_t)(s.ukeys[k0] >> 32), b = (uint32_t)s.ukeys[k0];
	__syncthreads();
	pal_place(s, a, 0, true);
	__syncthreads();
	pal_place(s, b, 1, true);
	__syncthreads();
	int lo = 0, hi = 2; // placed list = virtual positions [lo, hi)

	for (uint32_t step = 2; step < n; step++)
	{
		// find_next_entry (enc.cpp:1868): largest total, first unplaced entry on ties (the to-do list stays in ascending order)
		unsigned long long best = 0; bool any = false;
		for (uint32_t u = threadIdx.x; u < n; u += blockDim.x)
			if (!s.placed[u])
			{
				const unsigned long long cand = ((unsigned long long)s.total[u] << 32) | (0xFFFFFFFFu - u);
				if (!any || cand > best) { best = cand; any = true; }
			}
		if (!any) best = 0;
		for (int o = 16; o; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o); best = t > best ? t : best; }
		if (!lane) red[warp] = best;
		__syncthreads();
		if (!threadIdx.x)
		{
			for (int w = 1; w < 32; w++) best = red[w] > best ? red[w] : best;
			const uint32_t e = 0xFFFFFFFFu - (uint32_t)best;
			// pick_side (enc.cpp:1893): which_side = sum over the placed list, in order, of float(r * count), r = P + 1 - 2 (j + 1)
			const int P = hi - lo;
			float which_side = 0.0f;
			const uint32_t r0 = s.row_ptr[e], r1 = s.row_ptr[e + 1], nf = s.n_front[e], nb = s.n_back[e];
			for (uint32_t q = r1 - nf; q < r1; q++) // fronts: most recently placed (smallest position) first
			{
				const int j = s.nb_vpos[q] - lo, r = P + 1 - 2 * (j + 1);
				which_side += (float)(int)((uint32_t)r * s.nb_count[q]); // int product as the reference computes it
			}
			for (uint32_t q = r0; q < r0 + nb; q++)
			{
				const int j = s.nb_vpos[q] - lo, r = P + 1 - 2 * (j + 1);
				which_side += (float)(int)((uint32_t)r * s.nb_count[q]);
			}
			sh_e = e; sh_back = (which_side <= 0.0f) ? 1 : 0;
			sh_vp = sh_back ? hi : lo - 1;
		}
		__syncthreads();
		const uint32_t e = sh_e; const int vp = sh_vp; const bool back = sh_back != 0;
		if (back) hi++; else lo--;
		pal_place(s, e, vp, back);
		__syncthreads();
	}
	// m_remap_table[m_entries_picked[i]] = i
	for (uint32_t u = threadIdx.x; u < n; u += blockDim.x) s.remap[u] = (uint32_t)(s.vpos[u] - lo);
	if (!threadIdx.x) *s.status = 1;
}

} // namespace

#define PAL_OK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { ctx->fail_cuda(#expr, e_); ok = false; goto done; } } while (0)

extern "C" int b200_palette_reorder(b200_context* ctx, uint32_t num_indices, const uint32_t* pIndices, uint32_t num_syms, uint32_t* pRemap_table)
{
	if (!ctx || !ctx->activate()) return 0;
	if (!num_syms || !num_indices || !pIndices || !pRemap_table) { ctx->fail("b200_palette_reorder: null or empty input"); return 0; }
	for (uint32_t i = 0; i < num_indices; i++) if (pIndices[i] >= num_syms) { ctx->fail("b200_palette_reorder: index out of range"); return 0; }
	for (uint32_t i = 0; i < num_syms; i++) pRemap_table[i] = 0;
	if (num_indices <= 1) return 1; // the reference returns with a zeroed table (enc.cpp:1797)
	if (num_syms == 1) return 1; // placed list (0, 0), and only i = 0 is written: remap[0] = 0

	const uint32_t m = num_indices, n = num_syms;
	const size_t npairs = 2 * (size_t)(m - 1);
	bool ok = true;
	cudaStream_t st = ctx->stream;
	uint32_t* d_idx = nullptr; unsigned long long *d_keys = nullptr, *d_keys2 = nullptr, *d_ukeys = nullptr; uint32_t *d_ucount = nullptr, *d_nruns = nullptr, *d_row = nullptr;
	uint32_t *d_total = nullptr, *d_nbc = nullptr, *d_nback = nullptr, *d_nfront = nullptr, *d_remap = nullptr, *d_status = nullptr; int *d_vpos = nullptr, *d_nbv = nullptr; uint8_t* d_placed = nullptr;
	void* d_tmp = nullptr; size_t tmp_sort = 0, tmp_rle = 0;
	uint32_t status = 0;
	pal_state s;
	PAL_OK(cudaMallocAsync(&d_idx, (size_t)m * 4, st));
	PAL_OK(cudaMallocAsync(&d_keys, npairs * 8, st)); PAL_OK(cudaMallocAsync(&d_keys2, npairs * 8, st));
	PAL_OK(cudaMallocAsync(&d_ukeys, npairs * 8, st)); PAL_OK(cudaMallocAsync(&d_ucount, npairs * 4, st));
	PAL_OK(cudaMallocAsync(&d_nruns, 8, st)); PAL_OK(cudaMallocAsync(&d_row, ((size_t)n + 2) * 4, st));
	PAL_OK(cudaMallocAsync(&d_total, (size_t)n * 4, st)); PAL_OK(cudaMallocAsync(&d_vpos, (size_t)n * 4, st)); PAL_OK(cudaMallocAsync(&d_placed, n, st));
	PAL_OK(cudaMallocAsync(&d_nbv, npairs * 4, st)); PAL_OK(cudaMallocAsync(&d_nbc, npairs * 4, st));
	PAL_OK(cudaMallocAsync(&d_nback, (size_t)n * 4, st)); PAL_OK(cudaMallocAsync(&d_nfront, (size_t)n * 4, st)); PAL_OK(cudaMallocAsync(&d_remap, (size_t)n * 4, st));
	d_status = d_nruns + 1;
	PAL_OK(cudaMemcpyAsync(d_idx, pIndices, (size_t)m * 4, cudaMemcpyHostToDevice, st));
	PAL_OK(cudaMemsetAsync(d_nruns, 0, 8, st));
	PAL_OK(cudaMemsetAsync(d_total, 0, (size_t)n * 4, st)); PAL_OK(cudaMemsetAsync(d_vpos, 0, (size_t)n * 4, st)); PAL_OK(cudaMemsetAsync(d_placed, 0, n, st));
	PAL_OK(cudaMemsetAsync(d_nback, 0, (size_t)n * 4, st)); PAL_OK(cudaMemsetAsync(d_nfront, 0, (size_t)n * 4, st)); PAL_OK(cudaMemsetAsync(d_remap, 0, (size_t)n * 4, st));
	PAL_OK(cub::DeviceRadixSort::SortKeys(nullptr, tmp_sort, d_keys, d_keys2, (int)npairs, 0, 64, st));
	PAL_OK(cub::DeviceRunLengthEncode::Encode(nullptr, tmp_rle, d_keys2, d_ukeys, d_ucount, d_nruns, (int)npairs, st));
	PAL_OK(cudaMallocAsync(&d_tmp, tmp_sort > tmp_rle ? tmp_sort : tmp_rle, st));
	ctx->launches = 0;
	PAL_OK(cudaEventRecord(ctx->ev0, st));
	k_pal_pairs<<<(m + 255) / 256, 256, 0, st>>>(d_idx, m, d_keys);
	PAL_OK(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_sort, d_keys, d_keys2, (int)npairs, 0, 64, st));
	PAL_OK(cub::DeviceRunLengthEncode::Encode(d_tmp, tmp_rle, d_keys2, d_ukeys, d_ucount, d_nruns, (int)npairs, st));
	k_pal_rows<<<(n + 1 + 255) / 256, 256, 0, st>>>(d_ukeys, d_nruns, d_row, n);
	s.ukeys = d_ukeys; s.ucount = d_ucount; s.row_ptr = d_row; s.total = d_total; s.vpos = d_vpos; s.placed = d_placed; s.nb_vpos = d_nbv; s.nb_count = d_nbc;
	s.n_back = d_nback; s.n_front = d_nfront; s.remap = d_remap; s.n_syms = n; s.status = d_status;
	k_pal_order<<<1, 1024, 0, st>>>(s, d_nruns);
	ctx->launches = 3; __atomic_add_fetch(&g_b200_total_launches, 3, __ATOMIC_RELAXED);
	PAL_OK(cudaGetLastError());
	PAL_OK(cudaEventRecord(ctx->ev1, st));
	PAL_OK(cudaMemcpyAsync(pRemap_table, d_remap, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
	PAL_OK(cudaMemcpyAsync(&status, d_status, 4, cudaMemcpyDeviceToHost, st));
	PAL_OK(cudaStreamSynchronize(st));
	PAL_OK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
	if (status == 2)
	{
		// no two different indices are adjacent: hist is all zero, find_initial picks (0, 0), the rest follow in ascending order
		// (enc.cpp:1845-1866): placed list 0, 0, 1, 2, ..., and m_remap_table[list[i]] = i for i < num_syms
		for (uint32_t i = 0; i < n; i++) pRemap_table[i] = 0;
		pRemap_table[0] = 1;
		for (uint32_t i = 2; i < n; i++) pRemap_table[i - 1] = i;
	}
	else if (status != 1) { ctx->fail("b200_palette_reorder: kernel did not finish"); ok = false; }
	if (ok) ctx->account(B200_STAT_ETC1S_BACKEND_PREDICTION);
done:
	cudaFreeAsync(d_idx, st); cudaFreeAsync(d_keys, st); cudaFreeAsync(d_keys2, st); cudaFreeAsync(d_ukeys, st); cudaFreeAsync(d_ucount, st); cudaFreeAsync(d_nruns, st); cudaFreeAsync(d_row, st);
	cudaFreeAsync(d_total, st); cudaFreeAsync(d_vpos, st); cudaFreeAsync(d_placed, st); cudaFreeAsync(d_nbv, st); cudaFreeAsync(d_nbc, st); cudaFreeAsync(d_nback, st); cudaFreeAsync(d_nfront, st);
	cudaFreeAsync(d_remap, st); cudaFreeAsync(d_tmp, st);
	return ok ? 1 : 0;
}
