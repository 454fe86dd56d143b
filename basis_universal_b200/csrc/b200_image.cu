// b200_image.cu -- the steps either side of the per-block encode, so that a raster
// image can go in and a quality figure can come out without a host pass over the texels:
//   k_extract_blocks   raster RGBA8 -> 64 B pixel_blocks, edge texels clamped (basis_compressor::extract_source_blocks,
//                      encoder/basisu_comp.cpp:3207 -> image::extract_block_clamped, encoder/basisu_enc.h:3168)
//   k_unpack_uastc     16 B UASTC blocks -> 64 B RGBA blocks (basist::unpack_uastc(blk, pPixels, srgb = false),
//                      transcoder/basisu_transcoder.cpp:15886)
//   k_block_metrics    |a - b| histograms of two block arrays over the texels inside the image: the integer part of
//                      image_metrics::calc (encoder/basisu_enc.cpp:2155); the host finishes with the reference's double arithmetic
// All three are HBM-bound streaming kernels (80 / 80 / 128 B per block).
#include "b200_internal.h"
#include "bu_rdo.h"

using namespace bu;

#include "b200_tables.cuh"

// One thread per block; interior blocks of 16-byte-aligned rasters take four 128-bit loads.
__global__ void __launch_bounds__(256) k_extract_blocks(const uint8_t* __restrict__ img, uint32_t width, uint32_t height, size_t pitch, uint32_t nbx, uint32_t nby, uint4* __restrict__ blocks)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nbx * nby) return;
	const uint32_t bx = i % nbx, by = i / nbx;
	uint4* dst = blocks + (size_t)i * 4;
	const bool interior = (bx * 4 + 4 <= width) && (by * 4 + 4 <= height);
	if (interior && ((pitch | (size_t)img) & 15) == 0)
	{
#pragma unroll
		for (uint32_t y = 0; y < 4; y++)
			dst[y] = __ldg(reinterpret_cast<const uint4*>(img + (size_t)(by * 4 + y) * pitch) + bx);
		return;
	}
	for (uint32_t y = 0; y < 4; y++)
	{
		const uint32_t sy = min(by * 4 + y, height - 1);
		uint32_t p[4];
		for (uint32_t x = 0; x < 4; x++)
		{
			const uint32_t sx = min(bx * 4 + x, width - 1);
			p[x] = __ldg(reinterpret_cast<const uint32_t*>(img + (size_t)sy * pitch) + sx);
		}
		dst[y] = make_uint4(p[0], p[1], p[2], p[3]);
	}
}

__global__ void __launch_bounds__(128) k_unpack_uastc(const uint4* __restrict__ ublocks, uint32_t n, uint4* __restrict__ out, int* failed)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint4 v = __ldg(ublocks + i);
	block_bits b; b.lo = v.x | ((uint64_t)v.y << 32); b.hi = v.z | ((uint64_t)v.w << 32);
	uint32_t px[16];
	if (!unpack_block_texels(&d_tables, b, px)) *failed = 1;
	uint4* dst = out + (size_t)i * 4;
#pragma unroll
	for (int r = 0; r < 4; r++) dst[r] = make_uint4(px[r * 4], px[r * 4 + 1], px[r * 4 + 2], px[r * 4 + 3]);
}

// ETC1 block -> 16 RGBA texels: basisu::unpack_etc1(block, pDst, preserve_alpha = false) (encoder/basisu_etc.cpp:604). Handles
// the whole ETC1 format (individual and differential colours, both flip orientations), not only the ETC1S subset.
__global__ void __launch_bounds__(256) k_unpack_etc1(const uint2* __restrict__ eblocks, uint32_t n, uint4* __restrict__ out, int* failed)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint2 w = __ldg(eblocks + i);
	const uint32_t b[3] = { w.x & 255u, (w.x >> 8) & 255u, (w.x >> 16) & 255u }, b3 = w.x >> 24;
	const bool diff = (b3 & 2u) != 0, flip = (b3 & 1u) != 0;
	const uint32_t table[2] = { b3 >> 5, (b3 >> 2) & 7u };
	int base[2][3];
	bool ok = true;
	for (int c = 0; c < 3; c++)
	{
		if (diff)
		{
			const int c5 = (int)(b[c] >> 3);
			int d = (int)(b[c] & 7u); if (d >= 4) d -= 8;
			int c5b = c5 + d;
			if ((unsigned)c5b > 31u) { ok = false; c5b = clampi(c5b, 0, 31); } // unpack_color5 (etc.cpp:413) reports this
			base[0][c] = (c5 << 3) | (c5 >> 2);
			base[1][c] = (c5b << 3) | (c5b >> 2);
		}
		else
		{
			const int c4a = (int)(b[c] >> 4), c4b = (int)(b[c] & 15u);
			base[0][c] = (c4a << 4) | c4a;
			base[1][c] = (c4b << 4) | c4b;
		}
	}
	if (!ok) *failed = 1;
	const uint32_t msb = ((w.y & 255u) << 8) | ((w.y >> 8) & 255u), lsb = (((w.y >> 16) & 255u) << 8) | (w.y >> 24);
	uint32_t px[16];
	for (uint32_t y = 0; y < 4; y++)
		for (uint32_t x = 0; x < 4; x++)
		{
			const uint32_t bit = x * 4 + y;
			const uint32_t raw = (((msb >> bit) & 1u) << 1) | ((lsb >> bit) & 1u);
			const uint32_t sel = (0x1Eu >> (raw * 2)) & 3u; // g_etc1_to_selector_index = { 2, 3, 1, 0 }
			const uint32_t sub = flip ? (y >> 1) : (x >> 1);
			const int m = d_tables.etc1_inten[table[sub] * 4 + sel];
			// an overflowing block: unpack_etc1 returns false before writing any texel (etc.cpp:620); zeros are written here
			px[x + y * 4] = ok ? px_make((uint32_t)clamp255i(base[sub][0] + m), (uint32_t)clamp255i(base[sub][1] + m), (uint32_t)clamp255i(base[sub][2] + m), 255u) : 0u;
		}
	uint4* dst = out + (size_t)i * 4;
#pragma unroll
	for (int r = 0; r < 4; r++) dst[r] = make_uint4(px[r * 4], px[r * 4 + 1], px[r * 4 + 2], px[r * 4 + 3]);
}

// hist[6][256]: R, G, B, A, 709 luma, 601 luma; sums[8]: per-channel sums of a then of b (image_metrics::m_sum_a / m_sum_b).
// CTA-private histograms in shared memory, flushed with one atomic per non-empty bin.
__global__ void __launch_bounds__(256) k_block_metrics(const uint4* __restrict__ a, const uint4* __restrict__ b, uint32_t nbx, uint32_t nby, uint32_t width, uint32_t height,
	unsigned long long* __restrict__ hist, unsigned long long* __restrict__ sums)
{
	__shared__ uint32_t sh[6 * 256];
	__shared__ unsigned long long ssum[8];
	for (uint32_t k = threadIdx.x; k < 6 * 256; k += blockDim.x) sh[k] = 0;
	if (threadIdx.x < 8) ssum[threadIdx.x] = 0;
	__syncthreads();
	uint32_t local_sum[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
	const uint32_t total_rows = nbx * nby * 4; // one thread per block row (4 texels)
	for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < total_rows; r += gridDim.x * blockDim.x)
	{
		const uint32_t blk = r >> 2, y = r & 3, bx = blk % nbx, by = blk / nbx;
		if (by * 4 + y >= height) continue;
		const uint4 va = __ldg(a + r), vb = __ldg(b + r);
		const uint32_t pa[4] = { va.x, va.y, va.z, va.w }, pb[4] = { vb.x, vb.y, vb.z, vb.w };
		for (uint32_t x = 0; x < 4; x++)
		{
			if (bx * 4 + x >= width) break;
			const uint32_t d = __vabsdiffu4(pa[x], pb[x]);
			for (uint32_t c = 0; c < 4; c++)
			{
				atomicAdd(&sh[c * 256 + px_c(d, c)], 1u);
				local_sum[c] += px_c(pa[x], c); local_sum[4 + c] += px_c(pb[x], c);
			}
			// color_rgba::get_709_luma / get_601_luma (enc.h:1051-1052)
			const int la709 = (int)((13938u * px_c(pa[x], 0) + 46869u * px_c(pa[x], 1) + 4729u * px_c(pa[x], 2) + 32768u) >> 16);
			const int lb709 = (int)((13938u * px_c(pb[x], 0) + 46869u * px_c(pb[x], 1) + 4729u * px_c(pb[x], 2) + 32768u) >> 16);
			const int la601 = (int)((19595u * px_c(pa[x], 0) + 38470u * px_c(pa[x], 1) + 7471u * px_c(pa[x], 2) + 32768u) >> 16);
			const int lb601 = (int)((19595u * px_c(pb[x], 0) + 38470u * px_c(pb[x], 1) + 7471u * px_c(pb[x], 2) + 32768u) >> 16);
			atomicAdd(&sh[4 * 256 + (uint32_t)iabsi(la709 - lb709)], 1u);
			atomicAdd(&sh[5 * 256 + (uint32_t)iabsi(la601 - lb601)], 1u);
		}
	}
	for (int c = 0; c < 8; c++) if (local_sum[c]) atomicAdd(&ssum[c], (unsigned long long)local_sum[c]);
	__syncthreads();
	for (uint32_t k = threadIdx.x; k < 6 * 256; k += blockDim.x) if (sh[k]) atomicAdd(hist + k, (unsigned long long)sh[k]);
	if (threadIdx.x < 8 && ssum[threadIdx.x]) atomicAdd(sums + threadIdx.x, ssum[threadIdx.x]);
}

// ---- C ABI ------------------------------------------------------------------------------------------------------------------

static bool timed_begin(b200_context* ctx) { ctx->launches = 0; return cudaEventRecord(ctx->ev0, ctx->stream) == cudaSuccess; }
static bool timed_end(b200_context* ctx)
{
	if (cudaEventRecord(ctx->ev1, ctx->stream) != cudaSuccess) return false;
	if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return false;
	return cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1) == cudaSuccess;
}
static void count_launch(b200_context* ctx) { ctx->launches++; __atomic_add_fetch(&g_b200_total_launches, 1, __ATOMIC_RELAXED); }

static bool raster_pitch_ok(b200_context* ctx, const char* who, uint32_t width, size_t pitch_bytes)
{
	if (pitch_bytes >= (size_t)width * 4 && !(pitch_bytes & 3)) return true;
	snprintf(ctx->err, sizeof(ctx->err), "%s: pitch must be a multiple of 4 and >= width * 4", who);
	return false;
}

extern "C" int b200_extract_source_blocks_device(b200_context* ctx, const void* dRGBA, uint32_t width, uint32_t height, size_t pitch_bytes, void* dBlocks)
{
	if (!ctx || !ctx->activate()) return 0;
	if (!width || !height) { ctx->last_ms = 0; return 1; }
	if (!raster_pitch_ok(ctx, "b200_extract_source_blocks", width, pitch_bytes)) return 0;
	const uint32_t nbx = (width + 3) / 4, nby = (height + 3) / 4;
	if (!timed_begin(ctx)) return 0;
	k_extract_blocks<<<(nbx * nby + 255) / 256, 256, 0, ctx->stream>>>(static_cast<const uint8_t*>(dRGBA), width, height, pitch_bytes, nbx, nby, static_cast<uint4*>(dBlocks));
	count_launch(ctx);
	B200_CUDA_OK(ctx, cudaGetLastError());
	if (!timed_end(ctx)) { ctx->fail("b200_extract_source_blocks: kernel failed"); return 0; }
	return 1;
}

extern "C" int b200_extract_source_blocks(b200_context* ctx, const void* pRGBA, uint32_t width, uint32_t height, size_t pitch_bytes, void* pBlocks)
{
	if (!ctx || !ctx->activate()) return 0;
	if (!width || !height) { ctx->last_ms = 0; return 1; }
	if (!pRGBA || !pBlocks) { ctx->fail("b200_extract_source_blocks: null buffer"); return 0; }
	const uint32_t nbx = (width + 3) / 4, nby = (height + 3) / 4;
	if (!raster_pitch_ok(ctx, "b200_extract_source_blocks", width, pitch_bytes)) return 0;
	// The last row only guarantees width * 4 valid bytes (a sub-rectangle of a larger image ends there).
	const size_t in_bytes = pitch_bytes * (height - 1) + (size_t)width * 4, out_bytes = (size_t)nbx * nby * 64;
	if (!ctx->reserve(ctx->d_aux[0], ctx->aux_cap[0], in_bytes)) return 0;
	if (!ctx->reserve(ctx->d_in, ctx->in_cap, out_bytes)) return 0;
	B200_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_aux[0], pRGBA, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
	if (!b200_extract_source_blocks_device(ctx, ctx->d_aux[0], width, height, pitch_bytes, ctx->d_in)) return 0;
	B200_CUDA_OK(ctx, cudaMemcpy(pBlocks, ctx->d_in, out_bytes, cudaMemcpyDeviceToHost));
	return 1;
}

extern "C" int b200_uastc_unpack_blocks_device(b200_context* ctx, const void* dUastc, uint32_t num_blocks, void* dRGBA_blocks)
{
	if (!ctx || !ctx->activate()) return 0;
	if (!num_blocks) { ctx->last_ms = 0; return 1; }
	if (!ctx->reserve(ctx->d_aux[5], ctx->aux_cap[5], 256)) return 0;
	int* d_failed = static_cast<int*>(ctx->d_aux[5]);
	B200_CUDA_OK(ctx, cudaMemsetAsync(d_failed, 0, sizeof(int), ctx->stream));
	if (!timed_begin(ctx)) return 0;
	k_unpack_uastc<<<(num_blocks + 127) / 128, 128, 0, ctx->stream>>>(static_cast<const uint4*>(dUastc), num_blocks, static_cast<uint4*>(dRGBA_blocks), d_failed);
	count_launch(ctx);
	B200_CUDA_OK(ctx, cudaGetLastError());
	if (!timed_end(ctx)) { ctx->fail("b200_uastc_unpack_blocks: kernel failed"); return 0; }
	int h_failed = 0;
	B200_CUDA_OK(ctx, cudaMemcpy(&h_failed, d_failed, sizeof(int), cudaMemcpyDeviceToHost));
	if (h_failed) { ctx->fail("b200_uastc_unpack_blocks: invalid UASTC block (unpack_uastc returned false)"); return 0; }
	return 1;
}

extern "C" int b200_uastc_unpack_blocks(b200_context* ctx, const void* pUastc, uint32_t num_blocks, void* pRGBA_blocks)
{
	if (!ctx || !ctx->activate()) return 0;
	if (!num_blocks) { ctx->last_ms = 0; return 1; }
	if (!pUastc || !pRGBA_blocks) { ctx->fail("b200_uastc_unpack_blocks: null buffer"); return 0; }
	if (!ctx->reserve(ctx->d_out, ctx->out_cap, (size_t)num_blocks * 16)) return 0;
	if (!ctx->reserve(ctx->d_in, ctx->in_cap, (size_t)num_blocks * 64)) return 0;
	B200_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_out, pUastc, (size_t)num_blocks * 16, cudaMemcpyHostToDevice, ctx->stream));
	if (!b200_uastc_unpack_blocks_device(ctx, ctx->d_out, num_blocks, ctx->d_in)) return 0;
	B200_CUDA_OK(ctx, cudaMemcpy(pRGBA_blocks, ctx->d_in, (size_t)num_blocks * 64, cudaMemcpyDeviceToHost));
	return 1;
}

extern "C" int b200_block_metrics_device(b200_context* ctx, const void* dBlocksA, const void* dBlocksB, uint32_t width, uint32_t height, b200_block_metrics* pOut)
{
	if (!ctx || !ctx->activate()) return 0;
	if (!pOut) { ctx->fail("b200_block_metrics: null output"); return 0; }
	memset(pOut, 0, sizeof(*pOut));
	if (!width || !height) { ctx->last_ms = 0; return 1; }
	const uint32_t nbx = (width + 3) / 4, nby = (height + 3) / 4;
	const size_t bytes = sizeof(unsigned long long) * (6 * 256 + 8);
	if (!ctx->reserve(ctx->d_aux[5], ctx->aux_cap[5], bytes)) return 0;
	unsigned long long* d = static_cast<unsigned long long*>(ctx->d_aux[5]);
	B200_CUDA_OK(ctx, cudaMemsetAsync(d, 0, bytes, ctx->stream));
	if (!timed_begin(ctx)) return 0;
	const uint32_t rows = nbx * nby * 4;
	uint32_t grid = (rows + 255) / 256;
	if (grid > 148 * 8) grid = 148 * 8; // grid-stride: one CTA-private histogram flush per resident CTA
	k_block_metrics<<<grid, 256, 0, ctx->stream>>>(static_cast<const uint4*>(dBlocksA), static_cast<const uint4*>(dBlocksB), nbx, nby, width, height, d, d + 6 * 256);
	count_launch(ctx);
	B200_CUDA_OK(ctx, cudaGetLastError());
	if (!timed_end(ctx)) { ctx->fail("b200_block_metrics: kernel failed"); return 0; }
	B200_CUDA_OK(ctx, cudaMemcpy(pOut, d, bytes, cudaMemcpyDeviceToHost));
	return 1;
}

extern "C" int b200_uastc_encode_image(b200_context* ctx, const void* pRGBA, uint32_t width, uint32_t height, size_t pitch_bytes, void* pOut, uint32_t flags)
{
	if (!ctx || !ctx->activate()) return 0;
	if (!width || !height) { ctx->last_ms = 0; ctx->launches = 0; return 1; }
	if (!pRGBA || !pOut) { ctx->fail("b200_uastc_encode_image: null buffer"); return 0; }
	const uint32_t nbx = (width + 3) / 4, nby = (height + 3) / 4, n = nbx * nby;
	if (!raster_pitch_ok(ctx, "b200_uastc_encode_image", width, pitch_bytes)) return 0;
	const size_t in_bytes = pitch_bytes * (height - 1) + (size_t)width * 4; // see b200_extract_source_blocks
	if (!ctx->reserve(ctx->d_aux[0], ctx->aux_cap[0], in_bytes)) return 0;
	if (!ctx->reserve(ctx->d_in, ctx->in_cap, (size_t)n * 64)) return 0;
	if (!ctx->reserve(ctx->d_out, ctx->out_cap, (size_t)n * 16)) return 0;
	B200_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_aux[0], pRGBA, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
	if (!b200_extract_source_blocks_device(ctx, ctx->d_aux[0], width, height, pitch_bytes, ctx->d_in)) return 0;
	const float extract_ms = ctx->last_ms;
	if (!b200_uastc_encode_blocks_device(ctx, ctx->d_in, n, ctx->d_out, flags)) return 0;
	ctx->last_ms += extract_ms;
	ctx->launches += 1; // the ingest kernel
	B200_CUDA_OK(ctx, cudaMemcpy(pOut, ctx->d_out, (size_t)n * 16, cudaMemcpyDeviceToHost));
	return 1;
}

extern "C" int b200_etc1_unpack_blocks_device(b200_context* ctx, const void* dEtc1, uint32_t num_blocks, void* dRGBA_blocks)
{
	if (!ctx || !ctx->activate()) return 0;
	if (!num_blocks) { ctx->last_ms = 0; return 1; }
	if (!ctx->reserve(ctx->d_aux[5], ctx->aux_cap[5], 256)) return 0;
	int* d_failed = static_cast<int*>(ctx->d_aux[5]);
	B200_CUDA_OK(ctx, cudaMemsetAsync(d_failed, 0, sizeof(int), ctx->stream));
	if (!timed_begin(ctx)) return 0;
	k_unpack_etc1<<<(num_blocks + 255) / 256, 256, 0, ctx->stream>>>(static_cast<const uint2*>(dEtc1), num_blocks, static_cast<uint4*>(dRGBA_blocks), d_failed);
	count_launch(ctx);
	B200_CUDA_OK(ctx, cudaGetLastError());
	if (!timed_end(ctx)) { ctx->fail("b200_etc1_unpack_blocks: kernel failed"); return 0; }
	int h_failed = 0;
	B200_CUDA_OK(ctx, cudaMemcpy(&h_failed, d_failed, sizeof(int), cudaMemcpyDeviceToHost));
	if (h_failed) { ctx->fail("b200_etc1_unpack_blocks: a differential base colour overflowed (unpack_etc1 returns false for that block); its texels were zeroed"); return 0; }
	return 1;
}

extern "C" int b200_etc1_unpack_blocks(b200_context* ctx, const void* pEtc1, uint32_t num_blocks, void* pRGBA_blocks)
{
	if (!ctx || !ctx->activate()) return 0;
	if (!num_blocks) { ctx->last_ms = 0; return 1; }
	if (!pEtc1 || !pRGBA_blocks) { ctx->fail("b200_etc1_unpack_blocks: null buffer"); return 0; }
	if (!ctx->reserve(ctx->d_out, ctx->out_cap, (size_t)num_blocks * 8)) return 0;
	if (!ctx->reserve(ctx->d_in, ctx->in_cap, (size_t)num_blocks * 64)) return 0;
	B200_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_out, pEtc1, (size_t)num_blocks * 8, cudaMemcpyHostToDevice, ctx->stream));
	const int ok = b200_etc1_unpack_blocks_device(ctx, ctx->d_out, num_blocks, ctx->d_in);
	B200_CUDA_OK(ctx, cudaMemcpy(pRGBA_blocks, ctx->d_in, (size_t)num_blocks * 64, cudaMemcpyDeviceToHost)); // valid blocks are decoded even then
	return ok;
}

// ---- mip generation: basisu::image_resample for 8-bit images (encoder/basisu_enc.cpp:1022-1171) over basisu::Resampler
// (encoder/basisu_resampler.cpp:350-460) ------------------------------------------------------------------------------------
// The reference filters separably with per-destination contributor lists (source index + float weight), one axis after the other,
// in float: along X every output is `total = 0; total += s_j * w_j` in list order; along Y the first term is a plain product and the
// rest are added in list order; the result is clamped to [0, 1] and converted to 8 bits (linear: (int)(255 v + .5), sRGB channels
// through an 8192-entry table). Which axis goes first follows the reference's multiply-count estimate (resampler.cpp:772-806).
// The kernels keep exactly those operation orders (the library is built with --fmad=false), so the output bytes are the reference's;
// the contributor lists and the two sRGB tables are inputs (the caller gets them from Resampler / srgb_to_linear themselves).
struct resample_args
{
	const uint32_t* offsets; const b200_resample_contrib* contribs; // contributor list of the axis being filtered
	uint32_t out_w, out_h;          // size of this pass's output
	uint32_t in_w;                  // row length of this pass's input (texels)
	size_t in_pitch, out_pitch;     // bytes between rows of an 8-bit input / output
	uint32_t first_comp, num_comps;
	const float* s2l; const uint8_t* l2s; // sRGB tables or null
};

template<bool SRC_U8> __device__ __forceinline__ float resample_fetch(const void* in, const resample_args& a, uint32_t x, uint32_t y, uint32_t comp_index)
{
	if (SRC_U8)
	{
		const uint32_t v = static_cast<const uint8_t*>(in)[(size_t)y * a.in_pitch + (size_t)x * 4 + comp_index];
		return (a.s2l && comp_index != 3) ? a.s2l[v] : (float)v * (1.0f / 255.0f);
	}
	return static_cast<const float*>(in)[((size_t)y * a.in_w + x) * 4 + comp_index];
}

template<bool DST_U8> __device__ __forceinline__ void resample_store(void* out, const resample_args& a, uint32_t x, uint32_t y, uint32_t comp_index, float v)
{
	if (!DST_U8) { static_cast<float*>(out)[((size_t)y * a.out_w + x) * 4 + comp_index] = v; return; }
	if (v < 0.0f) v = 0.0f; else if (v > 1.0f) v = 1.0f; // Resampler::clamp with m_lo = 0, m_hi = 1
	uint8_t b;
	if (!a.l2s || comp_index == 3)
	{
		const int j = (int)(255.0f * v + .5f);
		b = (uint8_t)(j < 0 ? 0 : (j > 255 ? 255 : j));
	}
	else
	{
		const int j = (int)(8191.0f * v + .5f);
		b = a.l2s[j < 0 ? 0 : (j > 8191 ? 8191 : j)];
	}
	static_cast<uint8_t*>(out)[(size_t)y * a.out_pitch + (size_t)x * 4 + comp_index] = b;
}

template<bool SRC_U8, bool DST_U8> __global__ void __launch_bounds__(256) k_resample_x(const void* __restrict__ in, void* __restrict__ out, resample_args a)
{
	const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
	if (x >= a.out_w) return;
	const uint32_t c0 = a.offsets[x], c1 = a.offsets[x + 1];
	for (uint32_t c = 0; c < a.num_comps; c++)
	{
		const uint32_t comp_index = a.first_comp + c;
		float total = 0.0f;
		for (uint32_t j = c0; j < c1; j++) total += resample_fetch<SRC_U8>(in, a, a.contribs[j].pixel, y, comp_index) * a.contribs[j].weight;
		resample_store<DST_U8>(out, a, x, y, comp_index, total);
	}
}

template<bool SRC_U8, bool DST_U8> __global__ void __launch_bounds__(256) k_resample_y(const void* __restrict__ in, void* __restrict__ out, resample_args a)
{
	const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
	if (x >= a.out_w) return;
	const uint32_t c0 = a.offsets[y], c1 = a.offsets[y + 1];
	for (uint32_t c = 0; c < a.num_comps; c++)
	{
		const uint32_t comp_index = a.first_comp + c;
		float total = 0.0f;
		for (uint32_t j = c0; j < c1; j++)
		{
			const float t = resample_fetch<SRC_U8>(in, a, x, a.contribs[j].pixel, comp_index) * a.contribs[j].weight;
			total = (j == c0) ? t : total + t; // scale_y_mov for the first contributor, scale_y_add after
		}
		resample_store<DST_U8>(out, a, x, y, comp_index, total);
	}
}

extern "C" int b200_image_resample_rgba8(b200_context* ctx, const void* pSrc, uint32_t src_w, uint32_t src_h, size_t src_pitch_bytes, void* pDst, uint32_t dst_w, uint32_t dst_h,
	size_t dst_pitch_bytes, const uint32_t* pClist_x_offsets, const b200_resample_contrib* pClist_x, const uint32_t* pClist_y_offsets, const b200_resample_contrib* pClist_y,
	uint32_t first_comp, uint32_t num_comps, const float* pSrgb_to_linear, const uint8_t* pLinear_to_srgb)
{
	if (!ctx || !ctx->activate()) return 0;
	if (!pSrc || !pDst || !src_w || !src_h || !dst_w || !dst_h || !pClist_x_offsets || !pClist_x || !pClist_y_offsets || !pClist_y) { ctx->fail("b200_image_resample_rgba8: null or empty input"); return 0; }
	if (!num_comps || first_comp + num_comps > 4) { ctx->fail("b200_image_resample_rgba8: bad component range"); return 0; }
	if ((pSrgb_to_linear != nullptr) != (pLinear_to_srgb != nullptr)) { ctx->fail("b200_image_resample_rgba8: both sRGB tables or neither"); return 0; }
	if (!raster_pitch_ok(ctx, "b200_image_resample_rgba8", src_w, src_pitch_bytes) || !raster_pitch_ok(ctx, "b200_image_resample_rgba8", dst_w, dst_pitch_bytes)) return 0;
	const uint32_t nx = pClist_x_offsets[dst_w], ny = pClist_y_offsets[dst_h];
	for (uint32_t i = 0; i < nx; i++) if (pClist_x[i].pixel >= src_w) { ctx->fail("b200_image_resample_rgba8: X contributor outside the source"); return 0; }
	for (uint32_t i = 0; i < ny; i++) if (pClist_y[i].pixel >= src_h) { ctx->fail("b200_image_resample_rgba8: Y contributor outside the source"); return 0; }

	// Axis order: the reference's multiply-count estimate (resampler.cpp:772-806), ints as there.
	const int x_ops = (int)nx, y_ops = (int)ny;
	const int xy_ops = x_ops * (int)src_h + (4 * y_ops * (int)dst_w) / 3, yx_ops = (4 * y_ops * (int)src_w) / 3 + x_ops * (int)dst_h;
	const bool delay_x = (xy_ops > yx_ops) || ((xy_ops == yx_ops) && (src_w < dst_w));

	const size_t src_bytes = src_pitch_bytes * (src_h - 1) + (size_t)src_w * 4, dst_bytes = dst_pitch_bytes * (dst_h - 1) + (size_t)dst_w * 4;
	const uint32_t mid_w = delay_x ? src_w : dst_w, mid_h = delay_x ? dst_h : src_h;
	const size_t list_bytes = ((size_t)dst_w + 1 + (size_t)dst_h + 1) * 4 + ((size_t)nx + ny) * sizeof(b200_resample_contrib);
	if (!ctx->reserve(ctx->d_aux[0], ctx->aux_cap[0], src_bytes) || !ctx->reserve(ctx->d_aux[1], ctx->aux_cap[1], dst_bytes) ||
		!ctx->reserve(ctx->d_aux[2], ctx->aux_cap[2], (size_t)mid_w * mid_h * 16) || !ctx->reserve(ctx->d_aux[4], ctx->aux_cap[4], list_bytes) ||
		!ctx->reserve(ctx->d_aux[5], ctx->aux_cap[5], 256 * 4 + 8192)) return 0;
	uint8_t* d_lists = static_cast<uint8_t*>(ctx->d_aux[4]);
	uint32_t* d_xo = reinterpret_cast<uint32_t*>(d_lists);
	uint32_t* d_yo = d_xo + dst_w + 1;
	b200_resample_contrib* d_xc = reinterpret_cast<b200_resample_contrib*>(d_yo + dst_h + 1);
	b200_resample_contrib* d_yc = d_xc + nx;
	B200_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_aux[0], pSrc, src_bytes, cudaMemcpyHostToDevice, ctx->stream));
	B200_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_aux[1], pDst, dst_bytes, cudaMemcpyHostToDevice, ctx->stream)); // channels outside [first, first + num) keep their bytes
	B200_CUDA_OK(ctx, cudaMemcpyAsync(d_xo, pClist_x_offsets, ((size_t)dst_w + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
	B200_CUDA_OK(ctx, cudaMemcpyAsync(d_yo, pClist_y_offsets, ((size_t)dst_h + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
	B200_CUDA_OK(ctx, cudaMemcpyAsync(d_xc, pClist_x, (size_t)nx * sizeof(b200_resample_contrib), cudaMemcpyHostToDevice, ctx->stream));
	B200_CUDA_OK(ctx, cudaMemcpyAsync(d_yc, pClist_y, (size_t)ny * sizeof(b200_resample_contrib), cudaMemcpyHostToDevice, ctx->stream));
	float* d_s2l = nullptr; uint8_t* d_l2s = nullptr;
	if (pSrgb_to_linear)
	{
		d_s2l = static_cast<float*>(ctx->d_aux[5]); d_l2s = reinterpret_cast<uint8_t*>(d_s2l + 256);
		B200_CUDA_OK(ctx, cudaMemcpyAsync(d_s2l, pSrgb_to_linear, 256 * 4, cudaMemcpyHostToDevice, ctx->stream));
		B200_CUDA_OK(ctx, cudaMemcpyAsync(d_l2s, pLinear_to_srgb, 8192, cudaMemcpyHostToDevice, ctx->stream));
	}
	ctx->launches = 0;
	if (!timed_begin(ctx)) return 0;
	resample_args a;
	a.first_comp = first_comp; a.num_comps = num_comps; a.s2l = d_s2l; a.l2s = d_l2s;
	a.in_pitch = src_pitch_bytes; a.out_pitch = dst_pitch_bytes;
	if (!delay_x)
	{
		a.offsets = d_xo; a.contribs = d_xc; a.out_w = dst_w; a.out_h = src_h; a.in_w = src_w;
		k_resample_x<true, false><<<dim3((dst_w + 255) / 256, src_h), 256, 0, ctx->stream>>>(ctx->d_aux[0], ctx->d_aux[2], a);
		count_launch(ctx);
		a.offsets = d_yo; a.contribs = d_yc; a.out_w = dst_w; a.out_h = dst_h; a.in_w = dst_w;
		k_resample_y<false, true><<<dim3((dst_w + 255) / 256, dst_h), 256, 0, ctx->stream>>>(ctx->d_aux[2], ctx->d_aux[1], a);
		count_launch(ctx);
	}
	else
	{
		a.offsets = d_yo; a.contribs = d_yc; a.out_w = src_w; a.out_h = dst_h; a.in_w = src_w;
		k_resample_y<true, false><<<dim3((src_w + 255) / 256, dst_h), 256, 0, ctx->stream>>>(ctx->d_aux[0], ctx->d_aux[2], a);
		count_launch(ctx);
		a.offsets = d_xo; a.contribs = d_xc; a.out_w = dst_w; a.out_h = dst_h; a.in_w = src_w;
		k_resample_x<false, true><<<dim3((dst_w + 255) / 256, dst_h), 256, 0, ctx->stream>>>(ctx->d_aux[2], ctx->d_aux[1], a);
		count_launch(ctx);
	}
	B200_CUDA_OK(ctx, cudaGetLastError());
	if (!timed_end(ctx)) { ctx->fail("b200_image_resample_rgba8: kernel failed"); return 0; }
	B200_CUDA_OK(ctx, cudaMemcpy(pDst, ctx->d_aux[1], dst_bytes, cudaMemcpyDeviceToHost));
	return 1;
}
