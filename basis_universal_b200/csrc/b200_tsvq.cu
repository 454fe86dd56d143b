// b200_tsvq.cu -- tree-structured vector quantisation on sm_100a: b200_tsvq_generate (include/basisu_b200.h), the device form of
// basisu::generate_hierarchical_codebook_threaded + tree_vector_quant (encoder/basisu_enc.h:1546-2354), which the ETC1S frontend
// uses to cluster block endpoints (vec6F, frontend.cpp:868) and block selectors (vec16F, frontend.cpp:2140).
//
// The reference grows the tree best-first: a priority queue hands out the leaf with the largest variance, the leaf is split
// along its principal axis, two children are refined by <= 6 k-means passes and pushed back; one thread, ~9 passes over the
// node's members per split, up to 16 128 dependent splits. Two facts make this a GPU algorithm:
//   * a node's split depends only on the node (its member list and centroid), never on the order splits are taken in;
//   * a child's variance is (up to rounding) never larger than its parent's, so the final tree is "every node whose variance is
//     among the max_size - 1 largest", and a node below the current (max_size - 1)-th largest known variance can never be taken.
// So the tree is expanded SPECULATIVELY in rounds: every known, unsplit node whose variance still qualifies is split in the
// same launch (one CTA per node; the node's member list is a contiguous segment that the CTA partitions in place, stably), and
// the host then REPLAYS the reference's priority queue over the known nodes (tsvq_replay: same heap code path, same node
// numbering, same leaf/parent retrieval order). If the replay reaches a node whose split is not known yet, another round runs.
// The replay is the source of truth; speculation only decides what is computed ahead of time (~tree depth rounds in total).
//
//   k_tsvq_gather_key / cub radix sort   LSD sort of the training vectors by component => lexicographic order (the std::map order
//                                        of enc.h:2228-2260); k_tsvq_heads + scan + k_tsvq_groups merge duplicates, summing weights
//   k_tsvq_root<D>                       prepare_root (enc.h:1696): weighted centroid and variance of a member segment
//   k_tsvq_split<D, NT>                  split_node (enc.h:1723): covariance -> power-iteration PCA (compute_pca_from_covar,
//                                        enc.h:606) -> prep_split (enc.h:1848) -> refine_split (enc.h:1962) -> stable partition
//
// Float sums are order-dependent, and node variances (a difference of nearly equal sums) decide which nodes get split, so every
// sum is accumulated by one thread in the reference's member order (see "serial sums" below): the clusterer reproduces the CPU's
// codebooks exactly; per-member arithmetic (projections, distances, side tests) repeats the reference's operations one for one
// on all threads. Parallelism: across accumulators within a node, across the nodes of a round, and across members for the
// per-member work. Built with --fmad=false like the rest of the library (the reference is compiled without FMA contraction).
#include "b200_internal.h"
#include <cub/cub.cuh>
#include <algorithm>
#include <vector>

namespace
{
	struct dev_buf // grow-only device buffer from the stream-ordered pool (see b200_context::reserve)
	{
		void* p = nullptr; size_t cap = 0;
		bool reserve(size_t bytes, cudaStream_t s)
		{
			if (bytes <= cap) return true;
			if (p) { cudaFreeAsync(p, s); p = nullptr; cap = 0; }
			const size_t want = bytes + bytes / 4 + 256;
			if (cudaMallocAsync(&p, want, s) != cudaSuccess) { p = nullptr; return false; }
			cap = want;
			return true;
		}
		void release(cudaStream_t s) { if (p) cudaFreeAsync(p, s); p = nullptr; cap = 0; }
		template<typename T> T* as() { return static_cast<T*>(p); }
	};

	struct tsvq_state
	{
		dev_buf raw, tv, tw, perm[2], keys[2], cub_tmp, heads, gid, gstart, uvec, uw64, uwf, members, tmp_members, side, nodes, frontier, results;
		std::vector<uint32_t> cl_off, cl_idx, pa_off, pa_idx;
		void release(cudaStream_t s)
		{
			dev_buf* all[] = { &raw, &tv, &tw, &perm[0], &perm[1], &keys[0], &keys[1], &cub_tmp, &heads, &gid, &gstart, &uvec, &uw64, &uwf, &members, &tmp_members, &side, &nodes, &frontier, &results };
			for (dev_buf* b : all) b->release(s);
		}
	};

	template<int D> struct node_rec { uint32_t start, count; unsigned long long weight; float origin[D]; };
	struct split_out { uint32_t ok, l_count; float l_var, r_var; };
	struct root_out { float var; };

	// ---- duplicate merging -------------------------------------------------------------------------------------------------

	__global__ void k_tsvq_unpack(const uint8_t* __restrict__ raw, uint32_t n, uint32_t dim, size_t stride, size_t wofs, float* __restrict__ tv, unsigned long long* __restrict__ tw, uint32_t* __restrict__ perm)
	{
		const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i >= n) return;
		const uint8_t* r = raw + (size_t)i * stride;
		for (uint32_t c = 0; c < dim; c++) tv[(size_t)i * dim + c] = *reinterpret_cast<const float*>(r + c * 4);
		tw[i] = *reinterpret_cast<const unsigned long long*>(r + wofs);
		perm[i] = i;
	}

	// float -> u32 whose unsigned order is the float order
	__device__ __forceinline__ uint32_t ordered_bits(float f) { const uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }

	__global__ void k_tsvq_gather_key(const float* __restrict__ tv, const uint32_t* __restrict__ perm, uint32_t n, uint32_t dim, uint32_t comp, uint32_t* __restrict__ keys)
	{
		const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i < n) keys[i] = ordered_bits(tv[(size_t)perm[i] * dim + comp]);
	}

	__global__ void k_tsvq_heads(const float* __restrict__ tv, const uint32_t* __restrict__ perm, uint32_t n, uint32_t dim, uint32_t* __restrict__ heads)
	{
		const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i >= n) return;
		uint32_t h = 1;
		if (i)
		{
			const float* a = tv + (size_t)perm[i] * dim; const float* b = tv + (size_t)perm[i - 1] * dim;
			h = 0;
			for (uint32_t c = 0; c < dim; c++) if (a[c] != b[c]) { h = 1; break; }
		}
		heads[i] = h;
	}

	// gid = inclusive scan of heads (1-based group id). Group g's first sorted position, vector and summed weight.
	__global__ void k_tsvq_groups(const float* __restrict__ tv, const unsigned long long* __restrict__ tw, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ heads,
		const uint32_t* __restrict__ gid, uint32_t n, uint32_t dim, uint32_t* __restrict__ gstart, float* __restrict__ uvec, unsigned long long* __restrict__ uw64)
	{
		const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i >= n) return;
		const uint32_t g = gid[i] - 1, t = perm[i];
		if (heads[i])
		{
			gstart[g] = i;
			for (uint32_t c = 0; c < dim; c++) uvec[(size_t)g * dim + c] = tv[(size_t)t * dim + c];
		}
		atomicAdd(uw64 + g, tw[t]); // integer: order-free
		if (i == n - 1) gstart[g + 1] = n;
	}

	__global__ void k_tsvq_init_members(const unsigned long long* __restrict__ uw64, uint32_t u, float* __restrict__ uwf, uint32_t* __restrict__ members)
	{
		const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i >= u) return;
		uwf[i] = (float)uw64[i];
		members[i] = i;
	}

	// ---- block-wide sums in a fixed order ----------------------------------------------------------------------------------

	__device__ __forceinline__ double warp_sum_d(double v)
	{
#pragma unroll
		for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
		return v;
	}

	// Every thread contributes get(k), k < K; afterwards s_out[k] holds the sum (visible to all threads).
	template<int K, int NT, typename F> __device__ __forceinline__ void block_sum(F get, double* s_part, double* s_out)
	{
		const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
		for (int k = 0; k < K; k++)
		{
			const double x = warp_sum_d((double)get(k));
			if (!lane) s_part[warp * K + k] = x;
		}
		__syncthreads();
		for (int k = threadIdx.x; k < K; k += NT)
		{
			double a = 0;
			for (int w = 0; w < NT / 32; w++) a += s_part[w * K + k];
			s_out[k] = a;
		}
		__syncthreads();
	}

	template<int D> __device__ __forceinline__ void load_vec(const float* __restrict__ vecs, uint32_t m, float* v)
	{
		if (D % 4 == 0)
		{
			const float4* p = reinterpret_cast<const float4*>(vecs + (size_t)m * D);
#pragma unroll
			for (int q = 0; q < D / 4; q++) { const float4 t = __ldg(p + q); v[q * 4] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w; }
		}
		else
		{
			const float2* p = reinterpret_cast<const float2*>(vecs + (size_t)m * D);
#pragma unroll
			for (int q = 0; q < D / 2; q++) { const float2 t = __ldg(p + q); v[q * 2] = t.x; v[q * 2 + 1] = t.y; }
		}
	}

	template<int D> __device__ __forceinline__ float dot_f(const float* a, const float* b) // vec::dot_product (enc.h:473)
	{
		float r = a[0] * b[0];
#pragma unroll
		for (int i = 1; i < D; i++) r += a[i] * b[i];
		return r;
	}

	// ---- serial sums in the reference's order -------------------------------------------------------------------------------------
	// The reference accumulates a node's sums in float, one member after the other. A float sum's value depends on that order, and the
	// node variances computed from those sums (a difference of two nearly equal numbers) decide which nodes the priority queue splits.
	// So every sum that feeds a stored value or a decision is accumulated by ONE thread in member order, exactly as the CPU does it.
	// A pass over a node's members goes tile by tile (NT members): in the LOAD phase every thread takes one member, does the per-member
	// arithmetic (difference to the origin, projection, distances, side test, products) and stages the terms in shared memory; in the
	// ACCUMULATE phase one thread per accumulator (vector component / matrix entry: up to 136) adds the tile's terms in order, reading
	// shared memory only (the add chain is the critical path, ~NT x 4 cycles per tile). Parallelism: across accumulators within a
	// node, across the nodes of a round, across members for the per-member work.

	template<int D, int NT> struct tile_smem
	{
		float p[NT][D + 1]; // per member: the D terms (row padded to D + 1 words: conflict-free column writes)
		float t[NT];        // per member: a scalar term (weight, or the ttsum term)
		uint8_t side[NT];
	};

	// ACCUMULATE-phase helpers: add the tile's terms in member order; four members per iteration so that the shared-memory loads
	// run ahead of the add chain, which is the only dependency between iterations.
	template<int D, int NT> __device__ __forceinline__ float tile_sum_side(const tile_smem<D, NT>& S, uint32_t nt, uint32_t c, uint32_t want, float acc)
	{
		uint32_t k = 0;
		for (; k + 4 <= nt; k += 4)
		{
			const uint32_t sd = *reinterpret_cast<const uint32_t*>(S.side + k);
			const float a0 = S.p[k][c], a1 = S.p[k + 1][c], a2 = S.p[k + 2][c], a3 = S.p[k + 3][c];
			if ((sd & 0xFFu) == want) acc += a0;
			if (((sd >> 8) & 0xFFu) == want) acc += a1;
			if (((sd >> 16) & 0xFFu) == want) acc += a2;
			if ((sd >> 24) == want) acc += a3;
		}
		for (; k < nt; k++) if (S.side[k] == want) acc += S.p[k][c];
		return acc;
	}
	template<int D, int NT> __device__ __forceinline__ double tile_sum_side_t(const tile_smem<D, NT>& S, uint32_t nt, uint32_t want, double acc)
	{
		uint32_t k = 0;
		for (; k + 4 <= nt; k += 4)
		{
			const uint32_t sd = *reinterpret_cast<const uint32_t*>(S.side + k);
			const float a0 = S.t[k], a1 = S.t[k + 1], a2 = S.t[k + 2], a3 = S.t[k + 3];
			if ((sd & 0xFFu) == want) acc += (double)a0;
			if (((sd >> 8) & 0xFFu) == want) acc += (double)a1;
			if (((sd >> 16) & 0xFFu) == want) acc += (double)a2;
			if ((sd >> 24) == want) acc += (double)a3;
		}
		for (; k < nt; k++) if (S.side[k] == want) acc += (double)S.t[k];
		return acc;
	}

	// Pass driver. Plain form (PIPE = false, TS threads): load tile / barrier / accumulate / barrier. Pipelined form (PIPE = true,
	// 2 * TS threads, used for nodes with many members): threads TS .. 2TS-1 are LOADERS that stage tile t + 1 into the other buffer
	// while threads 0 .. TS-1 (the ACCUMULATORS among them) add tile t: the per-member arithmetic and its global loads run under the
	// add chain, one barrier per tile. load(member index, tile, slot); acc(tile, members in tile).
	template<bool PIPE, int TS, typename Tile, typename LoadF, typename AccF> __device__ __forceinline__ void run_pass(uint32_t count, Tile* tiles, LoadF load, AccF acc)
	{
		const uint32_t tid = threadIdx.x;
		if (!PIPE)
		{
			for (uint32_t base = 0; base < count; base += TS)
			{
				if (base + tid < count) load(base + tid, tiles[0], tid);
				__syncthreads();
				acc(tiles[0], (count - base < (uint32_t)TS) ? count - base : (uint32_t)TS);
				__syncthreads();
			}
		}
		else
		{
			const bool loader = tid >= (uint32_t)TS;
			const uint32_t slot = tid - TS;
			if (loader && slot < count) load(slot, tiles[0], slot);
			__syncthreads();
			uint32_t t = 0;
			for (uint32_t base = 0; base < count; base += TS, t ^= 1)
			{
				if (loader) { const uint32_t i = base + TS + slot; if (i < count) load(i, tiles[t ^ 1], slot); }
				else acc(tiles[t], (count - base < (uint32_t)TS) ? count - base : (uint32_t)TS);
				__syncthreads();
			}
		}
	}

	// ---- prepare_root (enc.h:1696-1721) -------------------------------------------------------------------------------------

	template<int D, int TS, bool PIPE> __global__ void __launch_bounds__(PIPE ? 2 * TS : TS) k_tsvq_root(const float* __restrict__ vecs, const float* __restrict__ wf, const unsigned long long* __restrict__ w64,
		const uint32_t* __restrict__ members, node_rec<D>* nodes, const uint32_t* __restrict__ root_ids, root_out* out)
	{
		constexpr int NT = PIPE ? 2 * TS : TS;
		__shared__ double s_part[(NT / 32) * 1];
		__shared__ double s_out[1];
		__shared__ tile_smem<D, TS> S[PIPE ? 2 : 1];
		__shared__ float s_sum[D];
		__shared__ double s_tt;
		node_rec<D>& nd = nodes[root_ids[blockIdx.x]];
		const uint32_t start = nd.start, count = nd.count, tid = threadIdx.x;
		double wsum = 0;
		float acc = 0.0f;
		double acc_d = 0.0;
		run_pass<PIPE, TS>(count, S,
			[&](uint32_t i, tile_smem<D, TS>& T, uint32_t slot)
			{
				const uint32_t m = members[start + i];
				float v[D];
				load_vec<D>(vecs, m, v);
				const float w = wf[m];
#pragma unroll
				for (int c = 0; c < D; c++) T.p[slot][c] = v[c] * w;   // root.m_origin += v * (float)weight
				T.t[slot] = dot_f<D>(v, v) * w;                          // ttsum += v.dot(v) * weight
				wsum += (double)w64[m];                                    // integers: exact in any order
			},
			[&](const tile_smem<D, TS>& T, uint32_t nt)
			{
				if (tid < D)
				{
					uint32_t k = 0;
					for (; k + 4 <= nt; k += 4) { const float a0 = T.p[k][tid], a1 = T.p[k + 1][tid], a2 = T.p[k + 2][tid], a3 = T.p[k + 3][tid]; acc += a0; acc += a1; acc += a2; acc += a3; }
					for (; k < nt; k++) acc += T.p[k][tid];
				}
				else if (tid == D)
				{
					uint32_t k = 0;
					for (; k + 4 <= nt; k += 4) { const float a0 = T.t[k], a1 = T.t[k + 1], a2 = T.t[k + 2], a3 = T.t[k + 3]; acc_d += (double)a0; acc_d += (double)a1; acc_d += (double)a2; acc_d += (double)a3; }
					for (; k < nt; k++) acc_d += (double)T.t[k];
				}
			});
		if (tid < D) s_sum[tid] = acc; else if (tid == D) s_tt = acc_d;
		block_sum<1, NT>([&](int) -> double { return wsum; }, s_part, s_out);
		if (!tid)
		{
			float org[D];
			for (int c = 0; c < D; c++) org[c] = s_sum[c];
			const unsigned long long weight = (unsigned long long)s_out[0];
			const float var = (float)(s_tt - (double)(dot_f<D>(org, org) / (float)weight));
			const float inv = 1.0f / (float)weight;
			for (int c = 0; c < D; c++) nd.origin[c] = org[c] * inv;
			nd.weight = weight;
			out[blockIdx.x].var = var;
		}
	}

	// ---- split_node ----------------------------------------------------------------------------------------------------------

	template<int D> __device__ void pca_axis(float* cm, float renorm, float* axis_out) // compute_split_axis tail + compute_pca_from_covar (enc.h:1831-1845, 606-649)
	{
		for (int x = 0; x < D; x++) for (int y = x; y < D; y++) cm[x * D + y] *= renorm;
		for (int x = 0; x < D - 1; x++) for (int y = x + 1; y < D; y++) cm[y * D + x] = cm[x * D + y];
		float axis[D], prev[D];
		for (int i = 0; i < D; i++) { const float t = i * (1.0f / (D - 1)); axis[i] = .75f + (1.25f - .75f) * t; prev[i] = axis[i]; } // lerp(a, b, s) = a + (b - a) * s
		for (int iter = 0; iter < 8; iter++)
		{
			float trial[D];
			double max_sum = 0;
			for (int i = 0; i < D; i++)
			{
				double sum = 0;
				for (int j = 0; j < D; j++) sum += (double)(cm[i * D + j] * axis[j]);
				trial[i] = (float)sum;
				max_sum = fmax(fabs(sum), max_sum);
			}
			if (max_sum != 0.0) { const float s = (float)(1.0f / max_sum); for (int i = 0; i < D; i++) trial[i] *= s; }
			float delta[D];
			for (int i = 0; i < D; i++) delta[i] = prev[i] - trial[i];
			for (int i = 0; i < D; i++) { prev[i] = axis[i]; axis[i] = trial[i]; }
			if (dot_f<D>(delta, delta) < .0024f) break;
		}
		const float len = sqrtf(dot_f<D>(axis, axis));
		if (len != 0.0f) { const float s = 1.0f / len; for (int i = 0; i < D; i++) axis[i] *= s; }
		for (int i = 0; i < D; i++) axis_out[i] = axis[i];
	}

	template<int D, int TS, bool PIPE> __global__ void __launch_bounds__(PIPE ? 2 * TS : TS) k_tsvq_split(const float* __restrict__ vecs, const float* __restrict__ wf, const unsigned long long* __restrict__ w64,
		uint32_t* members, uint32_t* tmp, uint8_t* side, node_rec<D>* nodes, const uint32_t* __restrict__ frontier, uint32_t child_base, split_out* out)
	{
		constexpr int NT = PIPE ? 2 * TS : TS;
		constexpr int NE = D * (D + 1) / 2; // upper-triangle entries of the covariance matrix, one thread each
		static_assert(TS >= NE && TS >= 2 * D + 2, "one thread per accumulator");
		typedef tile_smem<D, TS> tile_t;
		__shared__ double s_part[(NT / 32) * 3];
		__shared__ double s_out[3];
		__shared__ tile_t S[PIPE ? 2 : 1];
		__shared__ float s_origin[D], s_axis[D], s_l[D], s_r[D], s_cov[D * D], s_sum[2 * D];
		__shared__ double s_tt[2];
		__shared__ uint32_t s_warp[NT / 32];

		const uint32_t tid = threadIdx.x;
		const node_rec<D> nd = nodes[frontier[blockIdx.x]];
		const uint32_t start = nd.start, count = nd.count;
		const uint32_t* mem = members + start;
		uint8_t* sd = side + start;
		if (tid < D) s_origin[tid] = nd.origin[tid];
		__syncthreads();

		// ---- prep_split (enc.h:1848-1960)
		if (count == 2)
		{
			if (tid < D) { s_l[tid] = vecs[(size_t)mem[0] * D + tid]; s_r[tid] = vecs[(size_t)mem[1] * D + tid]; }
			__syncthreads();
		}
		else
		{
			// covariance (compute_split_axis, enc.h:1811-1822; the SSE4.1 16x16 kernel computes the same per-entry sums): entry (x, y >= x)
			{
				int ex = 0, ey = 0;
				if (tid < NE) { int k = (int)tid; while (k >= D - ex) { k -= D - ex; ex++; } ey = ex + k; }
				float acc = 0.0f;
				run_pass<PIPE, TS>(count, S,
					[&](uint32_t i, tile_t& T, uint32_t slot)
					{
						const uint32_t m = mem[i];
						float v[D];
						load_vec<D>(vecs, m, v);
#pragma unroll
						for (int c = 0; c < D; c++) T.p[slot][c] = v[c] - s_origin[c];
						T.t[slot] = wf[m];
					},
					[&](const tile_t& T, uint32_t nt)
					{
						if (tid >= NE) return;
						uint32_t k = 0;
						for (; k + 4 <= nt; k += 4)
						{
							const float q0 = T.p[k][ex] * (T.t[k] * T.p[k][ey]), q1 = T.p[k + 1][ex] * (T.t[k + 1] * T.p[k + 1][ey]);
							const float q2 = T.p[k + 2][ex] * (T.t[k + 2] * T.p[k + 2][ey]), q3 = T.p[k + 3][ex] * (T.t[k + 3] * T.p[k + 3][ey]);
							acc = acc + q0; acc = acc + q1; acc = acc + q2; acc = acc + q3;
						}
						for (; k < nt; k++) acc = acc + T.p[k][ex] * (T.t[k] * T.p[k][ey]);
					});
				if (tid < NE) s_cov[ex * D + ey] = acc;
			}
			__syncthreads();
			if (!tid) pca_axis<D>(s_cov, 1.0f / (float)nd.weight, s_axis);
			__syncthreads();

			for (int mode = 0; mode < 2; mode++) // 0: side of the principal axis; 1: the reference's fallback (first half / second half of the member list)
			{
				double lw = 0, rw = 0;
				const uint32_t half = count / 2;
				float acc = 0.0f;
				run_pass<PIPE, TS>(count, S,
					[&](uint32_t i, tile_t& T, uint32_t slot)
					{
						const uint32_t m = mem[i];
						float v[D], dv[D];
						load_vec<D>(vecs, m, v);
						const float w = wf[m];
#pragma unroll
						for (int c = 0; c < D; c++) dv[c] = v[c] - s_origin[c];
						const bool right = mode ? (i >= half) : (dot_f<D>(dv, s_axis) >= 0.0f);
						T.side[slot] = right ? 1 : 0;
#pragma unroll
						for (int c = 0; c < D; c++) T.p[slot][c] = v[c] * w;
						if (right) rw += (double)w; else lw += (double)w; // double sums of float-valued integers: exact in any order
					},
					[&](const tile_t& T, uint32_t nt) { if (tid < 2 * D) acc = tile_sum_side<D, TS>(T, nt, tid % D, tid / D, acc); });
				if (tid < 2 * D) s_sum[tid] = acc;
				block_sum<2, NT>([&](int k) -> double { return k ? rw : lw; }, s_part, s_out);
				const double l_weight = s_out[0], r_weight = s_out[1];
				if (!(l_weight > 0.0 && r_weight > 0.0)) continue;
				if (tid < D) { s_l[tid] = s_sum[tid] * (float)(1.0f / l_weight); s_r[tid] = s_sum[D + tid] * (float)(1.0f / r_weight); }
				break;
			}
			__syncthreads();
		}

		// ---- refine_split (enc.h:1962-2095)
		float prev_total_variance = 1e+10f;
		float l_var = 0, r_var = 0;
		unsigned long long l_weight = 0, r_weight = 0;
		uint32_t l_count = 0;
		bool ok = true;
		for (int iter = 0; iter < 6; iter++)
		{
			for (int degenerate = 0; degenerate < 2; degenerate++)
			{
				double lw = 0, rw = 0, lc = 0;
				float acc = 0.0f;
				double acc_d = 0.0;
				run_pass<PIPE, TS>(count, S,
					[&](uint32_t i, tile_t& T, uint32_t slot)
					{
						const uint32_t m = mem[i];
						float v[D];
						load_vec<D>(vecs, m, v);
						bool right;
						if (!degenerate)
						{
							double ld = 0, rd = 0; // vec::squared_distance_d (enc.h:483)
#pragma unroll
							for (int c = 0; c < D; c++) { const double a = (double)s_l[c] - (double)v[c]; ld += a * a; const double b = (double)s_r[c] - (double)v[c]; rd += b * b; }
							right = ld >= rd;
						}
						else right = (i == 0); // members are unique vectors, so only the first one equals "firstVec" (enc.h:2037-2061)
						const float w = wf[m];
#pragma unroll
						for (int c = 0; c < D; c++) T.p[slot][c] = v[c] * w;  // v * (float)weight
						T.t[slot] = (float)w64[m] * dot_f<D>(v, v);             // weight * v.dot(v)
						T.side[slot] = right ? 1 : 0;
						sd[i] = right ? 1 : 0;
						if (right) rw += (double)w64[m]; else { lw += (double)w64[m]; lc += 1.0; }
					},
					[&](const tile_t& T, uint32_t nt)
					{
						if (tid < 2 * D) acc = tile_sum_side<D, TS>(T, nt, tid % D, tid / D, acc);
						else if (tid < 2 * D + 2) acc_d = tile_sum_side_t<D, TS>(T, nt, tid - 2 * D, acc_d);
					});
				if (tid < 2 * D) s_sum[tid] = acc; else if (tid < 2 * D + 2) s_tt[tid - 2 * D] = acc_d;
				block_sum<3, NT>([&](int k) -> double { return k == 0 ? lw : (k == 1 ? rw : lc); }, s_part, s_out);
				l_weight = (unsigned long long)s_out[0]; r_weight = (unsigned long long)s_out[1];
				l_count = (uint32_t)s_out[2];
				if (l_weight && r_weight) break;
			}
			if (!l_weight || !r_weight) { ok = false; break; }
			float nl[D], nr[D];
#pragma unroll
			for (int c = 0; c < D; c++) { nl[c] = s_sum[c]; nr[c] = s_sum[D + c]; }
			l_var = (float)(s_tt[0] - (double)(dot_f<D>(nl, nl) / (float)l_weight));
			r_var = (float)(s_tt[1] - (double)(dot_f<D>(nr, nr) / (float)r_weight));
			const float il = 1.0f / (float)l_weight, ir = 1.0f / (float)r_weight;
			__syncthreads(); // every thread has read the sums and the old centroids
			if (tid < D) { s_l[tid] = nl[tid] * il; s_r[tid] = nr[tid] * ir; }
			__syncthreads();
			const float total_var = l_var + r_var;
			if (total_var < .00001f) break;
			if (((prev_total_variance - total_var) / total_var) < .00125f) break;
			prev_total_variance = total_var;
		}

		if (!ok)
		{
			if (!tid) { split_out o; o.ok = 0; o.l_count = 0; o.l_var = o.r_var = 0; out[blockIdx.x] = o; }
			return;
		}

		// ---- stable partition of the member segment: left members first, both in their original order (l_children / r_children)
		__syncthreads();
		uint32_t lbase = 0, rbase = l_count;
		for (uint32_t c0 = 0; c0 < count; c0 += NT)
		{
			const uint32_t i = c0 + tid;
			const bool valid = i < count;
			const uint32_t m = valid ? mem[i] : 0;
			const bool left = valid && !sd[i];
			const uint32_t mask = __ballot_sync(0xffffffffu, left);
			const uint32_t lane = tid & 31, warp = tid >> 5;
			if (!lane) s_warp[warp] = __popc(mask);
			__syncthreads();
			uint32_t before = 0, total_l = 0;
			for (uint32_t w = 0; w < NT / 32; w++) { const uint32_t c = s_warp[w]; if (w < warp) before += c; total_l += c; }
			before += __popc(mask & ((1u << lane) - 1u));
			if (left) tmp[start + lbase + before] = m;
			else if (valid) tmp[start + rbase + (tid - before)] = m;
			const uint32_t in_chunk = (count - c0 < (uint32_t)NT) ? count - c0 : (uint32_t)NT;
			lbase += total_l; rbase += in_chunk - total_l;
			__syncthreads();
		}
		for (uint32_t i = tid; i < count; i += NT) members[start + i] = tmp[start + i];

		if (!tid)
		{
			const uint32_t r_count = count - l_count;
			// distinct members are distinct vectors, so a multi-member child with no variance gets the reference's 1e-4 (enc.h:1748-1774)
			if (l_var <= 0.0f && l_count > 1) l_var = 1e-4f;
			if (r_var <= 0.0f && r_count > 1) r_var = 1e-4f;
			node_rec<D>& L = nodes[child_base + 2 * blockIdx.x];
			node_rec<D>& R = nodes[child_base + 2 * blockIdx.x + 1];
			L.start = start; L.count = l_count; L.weight = l_weight;
			R.start = start + l_count; R.count = r_count; R.weight = r_weight;
			for (int c = 0; c < D; c++) { L.origin[c] = s_l[c]; R.origin[c] = s_r[c]; }
			split_out o; o.ok = 1; o.l_count = l_count; o.l_var = l_var; o.r_var = r_var;
			out[blockIdx.x] = o;
		}
	}

	// ---- host: the reference's priority queue and tree bookkeeping, replayed over the speculatively expanded nodes ----------------

	struct ref_heap // basisu::priority_queue (enc.h:1455-1544): same sift rules, hence the same order among equal priorities
	{
		struct entry { uint32_t index; float priority; };
		std::vector<entry> h; uint32_t size = 0;
		void init(uint32_t max_entries, uint32_t first_index, float first_priority) { h.assign((size_t)max_entries + 2, entry{ 0, 0 }); h[1] = entry{ first_index, first_priority }; size = 1; }
		uint32_t top_index() const { return h[1].index; }
		void delete_top()
		{
			h[1] = h[size]; size--;
			if (!size) return;
			const entry orig = h[1];
			uint32_t k = 1, c;
			while ((c = k << 1) <= size)
			{
				if (c < size && h[c].priority < h[c + 1].priority) ++c;
				if (orig.priority > h[c].priority) break;
				h[k] = h[c]; k = c;
			}
			h[k] = orig;
		}
		void add(uint32_t index, float priority)
		{
			size++;
			if (size >= h.size()) h.resize((size_t)size + 1);
			uint32_t k = size;
			for (;;)
			{
				const uint32_t p = k >> 1;
				if (!p || h[p].priority > priority) break;
				h[k] = h[p]; k = p;
			}
			h[k] = entry{ index, priority };
		}
	};

	enum { SPLIT_UNKNOWN = 0, SPLIT_OK = 1, SPLIT_FAILED = 2 };
	struct spec_node { float var; uint32_t start, count; int state; uint32_t left, right; }; // index == device node id
	struct replay_node { uint32_t spec; int left, right, codebook_index; };
	struct tree
	{
		uint32_t root; uint32_t max_size;
		std::vector<replay_node> nodes;
		std::vector<uint32_t> known; // spec ids discovered under this root
		bool done = false;
	};

	// tree_vector_quant::generate (enc.h:1631-1676) over the known nodes. Returns true when complete; otherwise `blocked` is the
	// node whose split must be computed next.
	bool tsvq_replay(tree& t, const std::vector<spec_node>& spec, uint32_t& blocked)
	{
		t.nodes.clear();
		t.nodes.push_back(replay_node{ t.root, -1, -1, -1 });
		ref_heap heap;
		heap.init(t.max_size, 0, spec[t.root].var);
		uint32_t leaves = 1, next_codebook_index = 0;
		while (heap.size && leaves < t.max_size)
		{
			const uint32_t ni = heap.top_index();
			heap.delete_top();
			const spec_node& s = spec[t.nodes[ni].spec];
			if (s.count <= 1) continue;
			if (s.state == SPLIT_UNKNOWN) { blocked = t.nodes[ni].spec; return false; }
			if (s.state == SPLIT_FAILED) continue;
			const uint32_t l = (uint32_t)t.nodes.size(), r = l + 1;
			t.nodes[ni].left = (int)l; t.nodes[ni].right = (int)r; t.nodes[ni].codebook_index = (int)next_codebook_index++;
			t.nodes.push_back(replay_node{ s.left, -1, -1, -1 });
			t.nodes.push_back(replay_node{ s.right, -1, -1, -1 });
			const spec_node& L = spec[s.left]; const spec_node& R = spec[s.right];
			if (L.var > 0.0f && L.count > 1) heap.add(l, L.var);
			if (R.var > 0.0f && R.count > 1) heap.add(r, R.var);
			leaves++;
		}
		return true;
	}

	struct segment { uint32_t start, count; };

	void retrieve_leaves(const tree& t, const std::vector<spec_node>& spec, std::vector<segment>& out) // tree_vector_quant::retrieve (enc.h:1573-1584)
	{
		for (const replay_node& n : t.nodes)
			if (n.left < 0) out.push_back(segment{ spec[n.spec].start, spec[n.spec].count });
	}

	void retrieve_parents(const tree& t, const std::vector<spec_node>& spec, uint32_t max_clusters, std::vector<segment>& out) // retrieve(max_clusters, ...) (enc.h:1599-1629)
	{
		std::vector<uint32_t> stack;
		uint32_t ni = 0;
		for (;;)
		{
			const replay_node& cur = t.nodes[ni];
			if (cur.left < 0 || (2 + cur.codebook_index) > (int)max_clusters)
			{
				out.push_back(segment{ spec[cur.spec].start, spec[cur.spec].count });
				if (stack.empty()) break;
				ni = stack.back(); stack.pop_back();
				continue;
			}
			stack.push_back((uint32_t)cur.right);
			ni = (uint32_t)cur.left;
		}
	}

	template<int D> struct tsvq_runner
	{
		b200_context* ctx; tsvq_state* st; uint32_t U;
		std::vector<spec_node> spec;
		uint32_t rounds = 0, splits = 0;

		bool ensure_nodes(size_t n)
		{
			const size_t bytes = n * sizeof(node_rec<D>);
			if (bytes <= st->nodes.cap) return true;
			dev_buf nb;
			if (!nb.reserve(bytes * 2, ctx->stream)) return false;
			if (st->nodes.p && !spec.empty() && cudaMemcpyAsync(nb.p, st->nodes.p, spec.size() * sizeof(node_rec<D>), cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) { nb.release(ctx->stream); return false; }
			if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { nb.release(ctx->stream); return false; }
			st->nodes.release(ctx->stream);
			st->nodes = nb;
			return true;
		}

		// Creates root nodes over the given member segments (prepare_root) and returns their spec ids.
		bool make_roots(const std::vector<segment>& segs, std::vector<uint32_t>& ids)
		{
			const uint32_t base = (uint32_t)spec.size(), n = (uint32_t)segs.size();
			if (!ensure_nodes((size_t)base + n)) { ctx->fail("b200_tsvq_generate: out of device memory (nodes)"); return false; }
			std::vector<node_rec<D>> recs(n);
			std::vector<uint32_t> rid(n);
			for (uint32_t i = 0; i < n; i++) { memset(&recs[i], 0, sizeof(recs[i])); recs[i].start = segs[i].start; recs[i].count = segs[i].count; rid[i] = base + i; }
			if (!st->frontier.reserve((size_t)n * 4, ctx->stream) || !st->results.reserve((size_t)n * sizeof(root_out), ctx->stream)) { ctx->fail("b200_tsvq_generate: out of device memory"); return false; }
			if (cudaMemcpyAsync(st->nodes.as<node_rec<D>>() + base, recs.data(), (size_t)n * sizeof(node_rec<D>), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return false;
			if (cudaMemcpyAsync(st->frontier.p, rid.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return false;
			k_tsvq_root<D, 256, true><<<n, 512, 0, ctx->stream>>>(st->uvec.as<float>(), st->uwf.as<float>(), st->uw64.as<unsigned long long>(), st->members.as<uint32_t>(),
				st->nodes.as<node_rec<D>>(), st->frontier.as<uint32_t>(), st->results.as<root_out>());
			ctx->launches++; __atomic_add_fetch(&g_b200_total_launches, 1, __ATOMIC_RELAXED);
			std::vector<root_out> ro(n);
			if (cudaMemcpyAsync(ro.data(), st->results.p, (size_t)n * sizeof(root_out), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return false;
			if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return false;
			ids.resize(n);
			for (uint32_t i = 0; i < n; i++) { spec.push_back(spec_node{ ro[i].var, segs[i].start, segs[i].count, SPLIT_UNKNOWN, 0, 0 }); ids[i] = base + i; }
			return true;
		}

		bool split_round(const std::vector<uint32_t>& frontier)
		{
			// big nodes first, one launch per CTA size class
			std::vector<uint32_t> order(frontier);
			std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return spec[a].count > spec[b].count; });
			const uint32_t n = (uint32_t)order.size();
			uint32_t n_big = 0;
			while (n_big < n && spec[order[n_big]].count > 256) n_big++;
			const uint32_t child_base = (uint32_t)spec.size();
			if (!ensure_nodes((size_t)child_base + 2 * (size_t)n)) { ctx->fail("b200_tsvq_generate: out of device memory (nodes)"); return false; }
			if (!st->frontier.reserve((size_t)n * 4, ctx->stream) || !st->results.reserve((size_t)n * sizeof(split_out), ctx->stream)) { ctx->fail("b200_tsvq_generate: out of device memory"); return false; }
			if (cudaMemcpyAsync(st->frontier.p, order.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return false;
			const float* vecs = st->uvec.as<float>(); const float* wf = st->uwf.as<float>(); const unsigned long long* w64 = st->uw64.as<unsigned long long>();
			uint32_t* members = st->members.as<uint32_t>(); uint32_t* tmp = st->tmp_members.as<uint32_t>(); uint8_t* side = st->side.as<uint8_t>();
			node_rec<D>* nodes = st->nodes.as<node_rec<D>>();
			constexpr int NT_SMALL = (D == 6) ? 64 : 160; // >= D (D + 1) / 2 accumulator threads
			if (n_big)
			{
				// many members: 256-member tiles, 256 accumulator-side threads + 256 loader threads (pipelined passes)
				k_tsvq_split<D, 256, true><<<n_big, 512, 0, ctx->stream>>>(vecs, wf, w64, members, tmp, side, nodes, st->frontier.as<uint32_t>(), child_base, st->results.as<split_out>());
				ctx->launches++; __atomic_add_fetch(&g_b200_total_launches, 1, __ATOMIC_RELAXED);
			}
			if (n > n_big)
			{
				k_tsvq_split<D, NT_SMALL, false><<<n - n_big, NT_SMALL, 0, ctx->stream>>>(vecs, wf, w64, members, tmp, side, nodes, st->frontier.as<uint32_t>() + n_big, child_base + 2 * n_big, st->results.as<split_out>() + n_big);
				ctx->launches++; __atomic_add_fetch(&g_b200_total_launches, 1, __ATOMIC_RELAXED);
			}
			std::vector<split_out> so(n);
			if (cudaMemcpyAsync(so.data(), st->results.p, (size_t)n * sizeof(split_out), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return false;
			const cudaError_t e = cudaStreamSynchronize(ctx->stream);
			if (e != cudaSuccess) { ctx->fail_cuda("b200_tsvq_generate: split kernel", e); return false; }
			spec.resize((size_t)child_base + 2 * (size_t)n, spec_node{ 0, 0, 0, SPLIT_FAILED, 0, 0 });
			for (uint32_t i = 0; i < n; i++)
			{
				spec_node& s = spec[order[i]];
				if (!so[i].ok) { s.state = SPLIT_FAILED; continue; }
				s.state = SPLIT_OK; s.left = child_base + 2 * i; s.right = s.left + 1;
				spec[s.left] = spec_node{ so[i].l_var, s.start, so[i].l_count, SPLIT_UNKNOWN, 0, 0 };
				spec[s.right] = spec_node{ so[i].r_var, s.start + so[i].l_count, s.count - so[i].l_count, SPLIT_UNKNOWN, 0, 0 };
			}
			rounds++; splits += n;
			return true;
		}

		// Grows every tree to its max_size (generate()), all trees advancing in the same device rounds.
		bool grow(std::vector<tree>& trees)
		{
			for (tree& t : trees) { t.known.assign(1, t.root); t.done = false; }
			for (;;)
			{
				std::vector<uint32_t> frontier;
				std::vector<float> vars;
				for (tree& t : trees)
				{
					if (t.done) continue;
					uint32_t blocked = 0;
					if (tsvq_replay(t, spec, blocked)) { t.done = true; continue; }
					// threshold: the (max_size - 1 + failed)-th largest variance among the known splittable nodes
					vars.clear();
					uint32_t failed = 0;
					for (uint32_t id : t.known) { const spec_node& s = spec[id]; if (s.count > 1) { vars.push_back(s.var); if (s.state == SPLIT_FAILED) failed++; } }
					float thresh = -1.0f;
					const size_t k = (size_t)t.max_size - 1 + failed + 8;
					if (vars.size() > k) { std::nth_element(vars.begin(), vars.begin() + k, vars.end(), [](float a, float b) { return a > b; }); thresh = vars[k]; }
					bool has_blocked = false;
					for (uint32_t id : t.known)
					{
						const spec_node& s = spec[id];
						if (s.state != SPLIT_UNKNOWN || s.count <= 1) continue;
						if (id == blocked) has_blocked = true;
						else if (!(s.var > 0.0f) || s.var < thresh) continue;
						frontier.push_back(id);
					}
					if (!has_blocked) frontier.push_back(blocked);
				}
				if (frontier.empty()) break;
				const uint32_t first_child = (uint32_t)spec.size();
				if (!split_round(frontier)) return false;
				// register the new children with their trees
				for (tree& t : trees)
				{
					if (t.done) continue;
					const size_t nk = t.known.size();
					for (size_t i = 0; i < nk; i++)
					{
						const spec_node& s = spec[t.known[i]];
						if (s.state == SPLIT_OK && s.left >= first_child) { t.known.push_back(s.left); t.known.push_back(s.right); }
					}
				}
			}
			return true;
		}
	};

	template<int D> int tsvq_run(b200_context* ctx, tsvq_state* st, uint32_t U, uint32_t max_codebook_size, uint32_t max_parent_codebook_size, uint32_t max_threads,
		std::vector<segment>& leaves, std::vector<segment>& parents, uint32_t& rounds, uint32_t& splits)
	{
		tsvq_runner<D> R; R.ctx = ctx; R.st = st; R.U = U;
		std::vector<uint32_t> ids;
		if (!R.make_roots(std::vector<segment>(1, segment{ 0, U }), ids)) return 0;

		// generate_hierarchical_codebook_threaded_internal (enc.h:2103-2214)
		const bool limit_clusterizers = U > max_codebook_size;
		if (U < 65536 * 4) max_threads = 1; // enc.h:2296
		const bool single = max_threads <= 1 || U < 256 || max_codebook_size < max_threads * 16;
		if (max_threads > 16) max_threads = 16; // cMaxThreads, applied after the test above as in the reference
		if (single)
		{
			std::vector<tree> T(1);
			T[0].root = ids[0]; T[0].max_size = max_codebook_size;
			if (!R.grow(T)) return 0;
			retrieve_leaves(T[0], R.spec, leaves);
			if (max_parent_codebook_size) retrieve_parents(T[0], R.spec, max_parent_codebook_size, parents);
		}
		else
		{
			std::vector<tree> T(1);
			T[0].root = ids[0]; T[0].max_size = max_threads;
			if (!R.grow(T)) return 0;
			std::vector<segment> initial;
			retrieve_leaves(T[0], R.spec, initial);
			if (initial.size() < max_threads)
			{
				leaves = initial;
				if (max_parent_codebook_size) retrieve_parents(T[0], R.spec, max_parent_codebook_size, parents);
			}
			else
			{
				std::vector<uint32_t> sub_ids;
				if (!R.make_roots(initial, sub_ids)) return 0;
				std::vector<tree> S(initial.size());
				for (size_t i = 0; i < S.size(); i++)
				{
					S[i].root = sub_ids[i];
					S[i].max_size = limit_clusterizers ? (max_codebook_size + max_threads - 1) / max_threads : initial[i].count;
				}
				if (!R.grow(S)) return 0;
				for (size_t i = 0; i < S.size(); i++) retrieve_leaves(S[i], R.spec, leaves);
				if (max_parent_codebook_size)
					for (size_t i = 0; i < S.size(); i++) retrieve_parents(S[i], R.spec, (max_parent_codebook_size + max_threads - 1) / max_threads, parents);
			}
		}
		rounds = R.rounds; splits = R.splits;
		return 1;
	}
} // namespace

void b200_tsvq_release(b200_context* ctx)
{
	if (!ctx || !ctx->tsvq) return;
	tsvq_state* st = static_cast<tsvq_state*>(ctx->tsvq);
	st->release(ctx->stream);
	cudaStreamSynchronize(ctx->stream);
	delete st;
	ctx->tsvq = nullptr;
}

#define TSVQ_OK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { ctx->fail_cuda("b200_tsvq_generate: " #expr, e_); return 0; } } while (0)

extern "C" int b200_tsvq_generate(b200_context* ctx, uint32_t dim, uint32_t num_training, const void* pTraining, size_t stride_bytes, size_t weight_offset_bytes,
	uint32_t max_codebook_size, uint32_t max_parent_codebook_size, uint32_t max_threads, int even_odd_input_pairs_equal, b200_tsvq_result* pResult)
{
	if (!ctx) return 0;
	if (!ctx->activate()) return 0;
	if (!pResult) { ctx->fail("b200_tsvq_generate: null result"); return 0; }
	memset(pResult, 0, sizeof(*pResult));
	if (dim != 6 && dim != 16) { ctx->fail("b200_tsvq_generate: dim must be 6 (vec6F) or 16 (vec16F)"); return 0; }
	if (!num_training || !pTraining || !max_codebook_size) { ctx->fail("b200_tsvq_generate: empty training set or codebook"); return 0; } // generate() returns false on an empty set
	if (stride_bytes < (size_t)dim * 4 || (stride_bytes & 3) || (weight_offset_bytes & 7) || weight_offset_bytes + 8 > stride_bytes) { ctx->fail("b200_tsvq_generate: bad record layout"); return 0; }
	// Pairs (2i, 2i+1) hold equal vectors in that mode (enc.h:2244-2262): merging them first or last gives the same groups,
	// and the stable sort keeps 2i ahead of 2i+1 inside a group, as the reference's push_back order does.
	(void)even_odd_input_pairs_equal;
	if (!ctx->tsvq) ctx->tsvq = new tsvq_state();
	tsvq_state* st = static_cast<tsvq_state*>(ctx->tsvq);
	const uint32_t n = num_training;
	ctx->launches = 0;

	if (!st->raw.reserve((size_t)n * stride_bytes, ctx->stream) || !st->tv.reserve((size_t)n * dim * 4, ctx->stream) || !st->tw.reserve((size_t)n * 8, ctx->stream) || !st->perm[0].reserve((size_t)n * 4, ctx->stream) || !st->perm[1].reserve((size_t)n * 4, ctx->stream) ||
		!st->keys[0].reserve((size_t)n * 4, ctx->stream) || !st->keys[1].reserve((size_t)n * 4, ctx->stream) || !st->heads.reserve((size_t)n * 4, ctx->stream) || !st->gid.reserve((size_t)n * 4, ctx->stream) || !st->gstart.reserve(((size_t)n + 1) * 4, ctx->stream) ||
		!st->uvec.reserve((size_t)n * dim * 4, ctx->stream) || !st->uw64.reserve((size_t)n * 8, ctx->stream) || !st->uwf.reserve((size_t)n * 4, ctx->stream) || !st->members.reserve((size_t)n * 4, ctx->stream) || !st->tmp_members.reserve((size_t)n * 4, ctx->stream) || !st->side.reserve(n, ctx->stream))
	{ ctx->fail("b200_tsvq_generate: out of device memory"); return 0; }

	TSVQ_OK(cudaEventRecord(ctx->ev0, ctx->stream));
	TSVQ_OK(cudaMemcpyAsync(st->raw.p, pTraining, (size_t)n * stride_bytes, cudaMemcpyHostToDevice, ctx->stream));
	const uint32_t grid = (n + 255) / 256;
	uint32_t launches = 0;
	k_tsvq_unpack<<<grid, 256, 0, ctx->stream>>>(st->raw.as<uint8_t>(), n, dim, stride_bytes, weight_offset_bytes, st->tv.as<float>(), st->tw.as<unsigned long long>(), st->perm[0].as<uint32_t>());
	launches++;

	// LSD radix sort by component dim-1 .. 0: stable, so the final order is lexicographic with ties in index order
	size_t tmp_bytes = 0;
	{
		cub::DoubleBuffer<uint32_t> dk(st->keys[0].as<uint32_t>(), st->keys[1].as<uint32_t>()), dv(st->perm[0].as<uint32_t>(), st->perm[1].as<uint32_t>());
		TSVQ_OK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, (int)n, 0, 32, ctx->stream));
		size_t scan_bytes = 0;
		TSVQ_OK(cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, st->heads.as<uint32_t>(), st->gid.as<uint32_t>(), (int)n, ctx->stream));
		if (scan_bytes > tmp_bytes) tmp_bytes = scan_bytes;
		if (!st->cub_tmp.reserve(tmp_bytes, ctx->stream)) { ctx->fail("b200_tsvq_generate: out of device memory"); return 0; }
	}
	int cur = 0; // which perm buffer holds the current permutation
	for (int c = (int)dim - 1; c >= 0; c--)
	{
		k_tsvq_gather_key<<<grid, 256, 0, ctx->stream>>>(st->tv.as<float>(), st->perm[cur].as<uint32_t>(), n, dim, (uint32_t)c, st->keys[0].as<uint32_t>());
		cub::DoubleBuffer<uint32_t> dk(st->keys[0].as<uint32_t>(), st->keys[1].as<uint32_t>()), dv(st->perm[cur].as<uint32_t>(), st->perm[cur ^ 1].as<uint32_t>());
		size_t tb = st->cub_tmp.cap;
		TSVQ_OK(cub::DeviceRadixSort::SortPairs(st->cub_tmp.p, tb, dk, dv, (int)n, 0, 32, ctx->stream));
		if (dv.Current() != st->perm[cur].as<uint32_t>()) cur ^= 1;
		launches += 2;
	}
	uint32_t* perm = st->perm[cur].as<uint32_t>();
	k_tsvq_heads<<<grid, 256, 0, ctx->stream>>>(st->tv.as<float>(), perm, n, dim, st->heads.as<uint32_t>());
	{
		size_t tb = st->cub_tmp.cap;
		TSVQ_OK(cub::DeviceScan::InclusiveSum(st->cub_tmp.p, tb, st->heads.as<uint32_t>(), st->gid.as<uint32_t>(), (int)n, ctx->stream));
	}
	TSVQ_OK(cudaMemsetAsync(st->uw64.p, 0, (size_t)n * 8, ctx->stream));
	k_tsvq_groups<<<grid, 256, 0, ctx->stream>>>(st->tv.as<float>(), st->tw.as<unsigned long long>(), perm, st->heads.as<uint32_t>(), st->gid.as<uint32_t>(), n, dim,
		st->gstart.as<uint32_t>(), st->uvec.as<float>(), st->uw64.as<unsigned long long>());
	launches += 3;
	uint32_t U = 0;
	TSVQ_OK(cudaMemcpyAsync(&U, st->gid.as<uint32_t>() + (n - 1), 4, cudaMemcpyDeviceToHost, ctx->stream));
	TSVQ_OK(cudaStreamSynchronize(ctx->stream));
	k_tsvq_init_members<<<(U + 255) / 256, 256, 0, ctx->stream>>>(st->uw64.as<unsigned long long>(), U, st->uwf.as<float>(), st->members.as<uint32_t>());
	launches++;
	TSVQ_OK(cudaGetLastError());
	__atomic_add_fetch(&g_b200_total_launches, launches, __ATOMIC_RELAXED);

	std::vector<segment> leaves, parents;
	uint32_t rounds = 0, splits = 0;
	const int ok = (dim == 6) ? tsvq_run<6>(ctx, st, U, max_codebook_size, max_parent_codebook_size, max_threads, leaves, parents, rounds, splits)
	                          : tsvq_run<16>(ctx, st, U, max_codebook_size, max_parent_codebook_size, max_threads, leaves, parents, rounds, splits);
	ctx->launches += launches;
	if (!ok) return 0;

	// expand clusters of unique vectors into clusters of training-vector indices (enc.h:2302-2338)
	std::vector<uint32_t> h_members(U), h_gstart((size_t)U + 1), h_perm(n);
	TSVQ_OK(cudaMemcpyAsync(h_members.data(), st->members.p, (size_t)U * 4, cudaMemcpyDeviceToHost, ctx->stream));
	TSVQ_OK(cudaMemcpyAsync(h_gstart.data(), st->gstart.p, ((size_t)U + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
	TSVQ_OK(cudaMemcpyAsync(h_perm.data(), perm, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	TSVQ_OK(cudaEventRecord(ctx->ev1, ctx->stream));
	TSVQ_OK(cudaStreamSynchronize(ctx->stream));
	TSVQ_OK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
	auto expand = [&](const std::vector<segment>& segs, std::vector<uint32_t>& off, std::vector<uint32_t>& idx)
	{
		off.clear(); idx.clear();
		off.reserve(segs.size() + 1); idx.reserve(n);
		off.push_back(0);
		for (const segment& s : segs)
		{
			// A node's member list is an ascending run of unique-vector ids in the reference (children are stable sub-lists of the root's
			// 0..U-1). Here a speculative split that the replay did not take has already partitioned its node's segment in place, so the
			// set is right but the order is [left | right]: restore it.
			std::sort(h_members.begin() + s.start, h_members.begin() + s.start + s.count);
			for (uint32_t i = 0; i < s.count; i++)
			{
				const uint32_t g = h_members[s.start + i];
				for (uint32_t j = h_gstart[g]; j < h_gstart[g + 1]; j++) idx.push_back(h_perm[j]);
			}
			off.push_back((uint32_t)idx.size());
		}
	};
	expand(leaves, st->cl_off, st->cl_idx);
	expand(parents, st->pa_off, st->pa_idx);
	pResult->num_unique = U;
	pResult->num_clusters = (uint32_t)leaves.size();
	pResult->cluster_offsets = st->cl_off.data(); pResult->cluster_indices = st->cl_idx.data();
	pResult->num_parent_clusters = (uint32_t)parents.size();
	pResult->parent_offsets = st->pa_off.data(); pResult->parent_indices = st->pa_idx.data();
	pResult->rounds = rounds; pResult->nodes_split = splits;
	ctx->account(B200_STAT_TSVQ);
	return 1;
}
