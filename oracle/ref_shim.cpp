// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Thin extern "C" veneer over the UNMODIFIED reference encoder, compiled from the sources where they
// lie under /root/reference by oracle/Makefile into oracle/_ref/libbasisu_ref.so.  It exposes the four
// hot-path seams of SURVEY.md section 8(b) with plain pointers so tests/ and bench.py's cpu_baseline /
// --impl reference legs can drive the reference through ctypes:
//
//   ref_encode_uastc_blocks  -> basisu::encode_uastc            (encoder/basisu_uastc_enc.h:68)
//   ref_uastc_rdo            -> basisu::uastc_rdo               (encoder/basisu_uastc_enc.h:139)
//   ref_compress_image       -> basisu::basis_compress          (encoder/basisu_comp.h:1259)
//   ref_etc1s_*              -> basisu_frontend / etc1_optimizer (encoder/basisu_frontend.h:115-156)
//
// Nothing in this file restates reference logic; it only marshals arguments.
#include "encoder/basisu_comp.h"
#include "encoder/basisu_enc.h"
#include "encoder/basisu_uastc_enc.h"
#include "encoder/basisu_bc7enc.h"
#include "encoder/basisu_etc.h"
#include "encoder/basisu_frontend.h"
#include "encoder/basisu_gpu_texture.h"
#include "encoder/basisu_opencl.h"
#include "encoder/basisu_resampler.h"
#include "transcoder/basisu_transcoder.h"

#include <atomic>
#include <thread>
#include <vector>
#include <cstring>

#define REF_API extern "C" __attribute__((visibility("default")))

using namespace basisu;

static std::atomic<bool> g_inited(false);

REF_API void ref_init()
{
	if (!g_inited.exchange(true))
		basisu_encoder_init(false, false);
}

// Encodes n 4x4 RGBA blocks (64 B each, [y][x], R first) to n UASTC blocks (16 B each) with `threads`
// host threads (block ranges are independent, so the split does not affect the bytes).
REF_API void ref_encode_uastc_blocks(const uint8_t* pBlocks, uint32_t n, uint8_t* pOut, uint32_t flags, uint32_t threads)
{
	ref_init();
	if (threads <= 1)
	{
		for (uint32_t i = 0; i < n; i++)
			encode_uastc(pBlocks + (size_t)i * 64, *reinterpret_cast<basist::uastc_block*>(pOut + (size_t)i * 16), flags);
		return;
	}
	std::atomic<uint32_t> next(0);
	const uint32_t chunk = 256; // same chunking as comp.cpp:2005
	std::vector<std::thread> pool;
	for (uint32_t t = 0; t < threads; t++)
		pool.emplace_back([&]() {
			for (;;)
			{
				const uint32_t first = next.fetch_add(chunk);
				if (first >= n) break;
				const uint32_t last = std::min(n, first + chunk);
				for (uint32_t i = first; i < last; i++)
					encode_uastc(pBlocks + (size_t)i * 64, *reinterpret_cast<basist::uastc_block*>(pOut + (size_t)i * 16), flags);
			}
		});
	for (auto& th : pool) th.join();
}

// In-place RDO post-pass. total_jobs has the meaning of uastc_enc.h:139 (comp.cpp:2078 passes min(4, threads)).
REF_API int ref_uastc_rdo(uint32_t n, uint8_t* pBlocks, const uint8_t* pBlock_pixels, float lambda, uint32_t dict_size,
	float max_allowed_rms_increase_ratio, float skip_block_rms_thresh, float max_smooth_block_std_dev, float smooth_block_max_error_scale,
	uint32_t flags, uint32_t total_jobs, uint32_t threads)
{
	ref_init();
	uastc_rdo_params p;
	p.m_lambda = lambda;
	p.m_lz_dict_size = dict_size;
	p.m_max_allowed_rms_increase_ratio = max_allowed_rms_increase_ratio;
	p.m_skip_block_rms_thresh = skip_block_rms_thresh;
	p.m_max_smooth_block_std_dev = max_smooth_block_std_dev;
	p.m_smooth_block_max_error_scale = smooth_block_max_error_scale;
	job_pool pool(threads ? threads : 1);
	return uastc_rdo(n, reinterpret_cast<basist::uastc_block*>(pBlocks), reinterpret_cast<const color_rgba*>(pBlock_pixels), p, flags, &pool, total_jobs) ? 1 : 0;
}

// Whole-image compress through the reference's public C-style API. Returns a malloc'd buffer (free with ref_free).
// fmt: 0 = ETC1S, 1 = UASTC LDR 4x4 (basist::basis_tex_format).
REF_API void* ref_compress_image(uint32_t fmt, const uint8_t* pRGBA, uint32_t w, uint32_t h, uint32_t flags_and_quality, float rdo_quality, size_t* pSize)
{
	ref_init();
	return basis_compress((basist::basis_tex_format)fmt, pRGBA, w, h, w, flags_and_quality, rdo_quality, pSize, nullptr);
}

REF_API void ref_free(void* p) { basis_free_data(p); }

// Decodes n UASTC blocks to RGBA (64 B per block) with the reference transcoder.
REF_API int ref_unpack_uastc_blocks(const uint8_t* pBlocks, uint32_t n, uint8_t* pPixels)
{
	ref_init();
	for (uint32_t i = 0; i < n; i++)
		if (!basist::unpack_uastc(*reinterpret_cast<const basist::uastc_block*>(pBlocks + (size_t)i * 16), reinterpret_cast<basist::color32*>(pPixels + (size_t)i * 64), false))
			return 0;
	return 1;
}

// ---- fine-grained hooks used by the per-function differential tests ------------------------------------------

// color_cell_compression (bc7enc.cpp:1364) in its UASTC configuration (mode 255, ASTC endpoint range, linear metric).
// weight_table: 1,2,3 = g_bc7_weights{1,2,3}, 4 = g_astc_weights4, 5 = g_astc_weights5.
REF_API uint64_t ref_color_cell_compression(const uint8_t* pPixels, uint32_t num_pixels, uint32_t weight_table, uint32_t endpoint_range, uint32_t has_alpha,
	uint32_t uber_level, uint32_t ls_passes, uint8_t* pLow4, uint8_t* pHigh4, uint8_t* pSelectors)
{
	ref_init();
	static const uint32_t* s_w[6] = { nullptr, basist::g_bc7_weights1, basist::g_bc7_weights2, basist::g_bc7_weights3, basist::g_astc_weights4, basist::g_astc_weights5 };
	static const float* s_wx[6] = { nullptr, g_bc7_weights1x, g_bc7_weights2x, g_bc7_weights3x, g_astc_weights4x, g_astc_weights5x };

	color_cell_compressor_params cp;
	memset(&cp, 0, sizeof(cp));
	cp.m_num_pixels = num_pixels;
	cp.m_pPixels = reinterpret_cast<const basist::color_quad_u8*>(pPixels);
	cp.m_num_selector_weights = 1u << weight_table;
	cp.m_pSelector_weights = s_w[weight_table];
	cp.m_pSelector_weightsx = reinterpret_cast<const bc7enc_vec4F*>(s_wx[weight_table]);
	cp.m_astc_endpoint_range = endpoint_range;
	cp.m_weights[0] = cp.m_weights[1] = cp.m_weights[2] = cp.m_weights[3] = 1;
	cp.m_has_alpha = has_alpha != 0;

	bc7enc_compress_block_params comp;
	memset(&comp, 0, sizeof(comp));
	comp.m_max_partitions_mode1 = 64;
	comp.m_least_squares_passes = ls_passes;
	comp.m_weights[0] = comp.m_weights[1] = comp.m_weights[2] = comp.m_weights[3] = 1;
	comp.m_uber_level = uber_level;

	color_cell_compressor_results res;
	uint8_t sel_temp[16];
	memset(&res, 0, sizeof(res));
	res.m_pSelectors = pSelectors;
	res.m_pSelectors_temp = sel_temp;

	const uint64_t err = color_cell_compression(255, &cp, &res, &comp);
	memcpy(pLow4, res.m_astc_low_endpoint.m_c, 4);
	memcpy(pHigh4, res.m_astc_high_endpoint.m_c, 4);
	return err;
}

// ---- ETC1S ---------------------------------------------------------------------------------------------------

// Per-block ETC1S encode exactly as basisu_frontend::init_etc1_images does on the CPU (frontend.cpp:733-822):
// etc1_optimizer with the quality the comp_level maps to, perceptual metric selectable, ETC1S constraint on.
REF_API void ref_etc1s_encode_blocks(const uint8_t* pBlocks, uint32_t n, uint8_t* pOut8, uint32_t perceptual, uint32_t comp_level)
{
	ref_init();
	for (uint32_t i = 0; i < n; i++)
	{
		etc1_optimizer optimizer;
		etc1_optimizer::params optimizer_params;
		etc1_optimizer::results optimizer_results;

		if (comp_level == 0)
			optimizer_params.m_quality = cETCQualityFast;
		else if (comp_level == 1)
			optimizer_params.m_quality = cETCQualityMedium;
		else if (comp_level == BASISU_MAX_ETC1S_COMPRESSION_LEVEL)
			optimizer_params.m_quality = cETCQualityUber;

		optimizer_params.m_num_src_pixels = 16;
		optimizer_params.m_pSrc_pixels = reinterpret_cast<const color_rgba*>(pBlocks + (size_t)i * 64);
		optimizer_params.m_perceptual = perceptual != 0;

		uint8_t selectors[16];
		optimizer_results.m_pSelectors = selectors;
		optimizer_results.m_n = 16;

		optimizer.init(optimizer_params, optimizer_results);
		optimizer.compute(); // the frontend treats a false return as an internal invariant failure (frontend.cpp:800)

		etc_block& blk = *reinterpret_cast<etc_block*>(pOut8 + (size_t)i * 8);
		memset(&blk, 0, sizeof(blk));
		blk.set_block_color5_etc1s(optimizer_results.m_block_color_unscaled);
		blk.set_inten_tables_etc1s(optimizer_results.m_block_inten_table);
		blk.set_flip_bit(true);

		for (uint32_t y = 0; y < 4; y++)
			for (uint32_t x = 0; x < 4; x++)
				blk.set_selector(x, y, selectors[x + y * 4]);
	}
}

// etc_block::determine_selectors (etc.h:374) given one (rgb5, inten) per block: the CPU meaning of opencl_determine_selectors.
REF_API void ref_etc1s_determine_selectors(const uint8_t* pBlocks, uint32_t n, const uint8_t* pRGB5_inten /*4 B per block*/, uint8_t* pOut8, uint32_t perceptual)
{
	ref_init();
	for (uint32_t i = 0; i < n; i++)
	{
		etc_block& blk = *reinterpret_cast<etc_block*>(pOut8 + (size_t)i * 8);
		memset(&blk, 0, sizeof(blk));
		const uint8_t* p = pRGB5_inten + (size_t)i * 4;
		blk.set_block_color5_etc1s(color_rgba(p[0], p[1], p[2], 255));
		blk.set_inten_tables_etc1s(p[3]);
		blk.set_flip_bit(true);
		blk.determine_selectors(reinterpret_cast<const color_rgba*>(pBlocks + (size_t)i * 64), perceptual != 0);
	}
}

// Decodes n ETC1 blocks with the reference (unpack_etc1, etc.h / etc.cpp).
REF_API void ref_unpack_etc1_blocks(const uint8_t* pBlocks8, uint32_t n, uint8_t* pPixels)
{
	ref_init();
	for (uint32_t i = 0; i < n; i++)
		unpack_etc1(*reinterpret_cast<const etc_block*>(pBlocks8 + (size_t)i * 8), reinterpret_cast<color_rgba*>(pPixels + (size_t)i * 64));
}

REF_API uint32_t ref_color_distance(uint32_t perceptual, const uint8_t* a, const uint8_t* b, uint32_t alpha)
{
	return color_distance(perceptual != 0, color_rgba(a[0], a[1], a[2], a[3]), color_rgba(b[0], b[1], b[2], b[3]), alpha != 0);
}

// Field dump of a packed UASTC block (basist::unpack_uastc, transcoder.cpp:15282) for test diagnostics:
// out[0]=mode [1]=pattern [2]=bc1_hint0 [3]=bc1_hint1 [4]=etc1 flip [5]=diff [6]=inten0 [7]=inten1 [8]=bias [9]=etc2 hints
// [10]=ccs [11..28]=endpoints [29..60]=weights
REF_API int ref_uastc_fields(const uint8_t* pBlock, uint8_t* out)
{
	ref_init();
	basist::unpacked_uastc_block u;
	memset(&u, 0, sizeof(u));
	if (!basist::unpack_uastc(*reinterpret_cast<const basist::uastc_block*>(pBlock), u, false, true))
		return 0;
	memset(out, 0, 61);
	out[0] = (uint8_t)u.m_mode; out[1] = (uint8_t)u.m_common_pattern; out[2] = u.m_bc1_hint0; out[3] = u.m_bc1_hint1;
	out[4] = u.m_etc1_flip; out[5] = u.m_etc1_diff; out[6] = (uint8_t)u.m_etc1_inten0; out[7] = (uint8_t)u.m_etc1_inten1;
	out[8] = (uint8_t)u.m_etc1_bias; out[9] = (uint8_t)u.m_etc2_hints;
	if (u.m_mode != 8)
	{
		out[10] = (uint8_t)u.m_astc.m_ccs;
		memcpy(out + 11, u.m_astc.m_endpoints, 18);
		memcpy(out + 29, u.m_astc.m_weights, 32);
	}
	else { out[11] = u.m_solid_color.r; out[12] = u.m_solid_color.g; out[13] = u.m_solid_color.b; out[14] = u.m_solid_color.a; out[15]=(uint8_t)u.m_etc1_r; out[16]=(uint8_t)u.m_etc1_g; out[17]=(uint8_t)u.m_etc1_b; out[18]=(uint8_t)u.m_etc1_selector; }
	return 1;
}

// ---- whole-file helpers for the end-to-end (drop-in) checks -----------------------------------------------------------------

// Initialises the encoder with its GPU seam enabled (basisu_encoder_init(use_opencl=true), enc.cpp:194-211). Returns
// opencl_is_available(): always 0 in the stock build (OpenCL compiled out), 1 in the drop-in build when a B200 is present.
REF_API int ref_init_gpu_seam()
{
	if (!g_inited.exchange(true))
		basisu_encoder_init(true, false);
	return opencl_is_available() ? 1 : 0;
}

// Transcodes image 0 / level 0 of a .basis file to RGBA32 with the reference transcoder (transcoder/basisu_transcoder.h).
REF_API int ref_transcode_basis_to_rgba(const void* pData, uint32_t size, uint8_t* pOut, uint32_t out_pixels)
{
	ref_init();
	basist::basisu_transcoder dec;
	if (!dec.validate_header(pData, size)) return 0;
	if (!dec.start_transcoding(pData, size)) return 0;
	basist::basisu_image_level_info li;
	if (!dec.get_image_level_info(pData, size, li, 0, 0)) return 0;
	if ((uint64_t)li.m_orig_width * li.m_orig_height > out_pixels) return 0;
	if (!dec.transcode_image_level(pData, size, 0, 0, pOut, li.m_orig_width * li.m_orig_height, basist::transcoder_texture_format::cTFRGBA32, 0, li.m_orig_width, nullptr, li.m_orig_height))
		return 0;
	return 1;
}

// The CPU path of basisu_frontend::generate_endpoint_codebook for one endpoint cluster (frontend.cpp:1523-1549):
// etc1_optimizer over `n` texels with the quality the comp_level maps to. out4 = {r5, g5, b5, inten}.
REF_API uint64_t ref_etc1s_encode_cluster(const uint8_t* pPixels, uint32_t n, uint32_t perceptual, uint32_t comp_level, uint8_t* out4)
{
	ref_init();
	etc1_optimizer optimizer;
	etc1_optimizer::params p;
	p.m_num_src_pixels = n;
	p.m_pSrc_pixels = reinterpret_cast<const color_rgba*>(pPixels);
	p.m_use_color4 = false;
	p.m_perceptual = perceptual != 0;
	if (comp_level <= 1) p.m_quality = cETCQualityMedium;
	else if (comp_level == BASISU_MAX_ETC1S_COMPRESSION_LEVEL) p.m_quality = cETCQualityUber;
	etc1_optimizer::results r;
	std::vector<uint8_t> sel(n);
	r.m_n = n;
	r.m_pSelectors = sel.data();
	optimizer.init(p, r);
	optimizer.compute();
	out4[0] = r.m_block_color_unscaled.r; out4[1] = r.m_block_color_unscaled.g; out4[2] = r.m_block_color_unscaled.b; out4[3] = (uint8_t)r.m_block_inten_table;
	return r.m_error;
}

// The per-cluster body of basisu_frontend::reoptimize_remapped_endpoints (frontend.cpp:3018-3090): `nblocks` source blocks
// (64 B each), one packed selector word per block (texel (x, y) at bits 2 * (x + 4 * y)), the cluster's current endpoint cur4 =
// {r5, g5, b5, inten}. Returns the optimiser's error; out4 = new endpoint, *pCur_err = error of cur4 with the imposed selectors.
REF_API uint64_t ref_etc1s_reoptimize_cluster(const uint8_t* pBlocks, uint32_t nblocks, const uint32_t* pSelectors, const uint8_t* cur4, uint32_t perceptual, uint32_t comp_level,
	uint8_t* out4, uint64_t* pCur_err)
{
	ref_init();
	const uint32_t n = nblocks * 16;
	std::vector<color_rgba> px(n);
	std::vector<uint8_t> force(n);
	etc_block blk;
	blk.set_block_color5_etc1s(color_rgba(cur4[0], cur4[1], cur4[2], 255));
	blk.set_inten_tables_etc1s(cur4[3]);
	blk.set_flip_bit(true);
	uint64_t cur_err = 0;
	for (uint32_t b = 0; b < nblocks; b++)
	{
		memcpy(&px[b * 16], pBlocks + (size_t)b * 64, 64);
		for (uint32_t y = 0; y < 4; y++)
			for (uint32_t x = 0; x < 4; x++)
			{
				const uint32_t sv = (pSelectors[b] >> (2 * (x + 4 * y))) & 3;
				force[b * 16 + x + y * 4] = (uint8_t)sv;
				blk.set_selector(x, y, sv);
			}
		cur_err += blk.evaluate_etc1_error(reinterpret_cast<const color_rgba*>(pBlocks + (size_t)b * 64), perceptual != 0);
	}
	etc1_optimizer optimizer;
	etc1_optimizer::params p;
	p.m_num_src_pixels = n;
	p.m_pSrc_pixels = px.data();
	p.m_use_color4 = false;
	p.m_perceptual = perceptual != 0;
	p.m_pForce_selectors = force.data();
	p.m_quality = (comp_level == BASISU_MAX_ETC1S_COMPRESSION_LEVEL) ? cETCQualityUber : cETCQualitySlow;
	etc1_optimizer::results r;
	std::vector<uint8_t> sel(n);
	r.m_n = n;
	r.m_pSelectors = sel.data();
	optimizer.init(p, r);
	if (!optimizer.compute()) { *pCur_err = cur_err; return UINT64_MAX; }
	out4[0] = r.m_block_color_unscaled.r; out4[1] = r.m_block_color_unscaled.g; out4[2] = r.m_block_color_unscaled.b; out4[3] = (uint8_t)r.m_block_inten_table;
	*pCur_err = cur_err;
	return r.m_error;
}

// The per-cluster body of generate_endpoint_codebook at step >= 1 (frontend.cpp:1493-1606): etc1_optimizer over the cluster's texels
// (quality by comp_level) and the error of the previous endpoint prev4 = {r5, g5, b5, inten}, best of four colours per texel.
REF_API uint64_t ref_etc1s_refit_cluster(const uint8_t* pBlocks, uint32_t nblocks, const uint8_t* prev4, uint32_t perceptual, uint32_t comp_level, uint8_t* out4, uint64_t* pPrev_err)
{
	const uint64_t new_err = ref_etc1s_encode_cluster(pBlocks, nblocks * 16, perceptual, comp_level, out4);
	color_rgba block_colors[4];
	etc_block::get_block_colors5(block_colors, color_rgba(prev4[0], prev4[1], prev4[2], 255), prev4[3], false);
	const color_rgba* px = reinterpret_cast<const color_rgba*>(pBlocks);
	uint64_t total = 0;
	for (uint32_t i = 0; i < nblocks * 16; i++)
	{
		uint64_t best = UINT64_MAX;
		for (uint32_t k = 0; k < 4; k++) best = minimum<uint64_t>(best, color_distance(perceptual != 0, px[i], block_colors[k], false));
		total += best;
	}
	*pPrev_err = total;
	return new_err;
}

// compute_endpoint_subblock_error_vec's inner body (frontend.cpp:1022-1066) for every block: pC5i = {r5, g5, b5, inten} per block.
REF_API void ref_etc1s_subblock_errors(const uint8_t* pBlocks, uint32_t n, const uint8_t* pC5i, uint32_t perceptual, uint64_t* pOut)
{
	ref_init();
	for (uint32_t b = 0; b < n; b++)
	{
		const color_rgba* px = reinterpret_cast<const color_rgba*>(pBlocks + (size_t)b * 64);
		color_rgba block_colors[4];
		etc_block::get_block_colors5(block_colors, color_rgba(pC5i[b * 4], pC5i[b * 4 + 1], pC5i[b * 4 + 2], 255), pC5i[b * 4 + 3], true);
		for (uint32_t sub = 0; sub < 2; sub++)
		{
			uint64_t total = 0;
			for (uint32_t i = 0; i < 8; i++)
			{
				const color_rgba& c = px[g_etc1_pixel_indices[1][sub][i]];
				uint64_t best = UINT64_MAX;
				for (uint32_t k = 0; k < 4; k++) best = minimum<uint64_t>(best, color_distance(perceptual != 0, c, block_colors[k], false));
				total += best;
			}
			pOut[b * 2 + sub] = total;
		}
	}
}

// basisu_backend::create_encoder_blocks' endpoint-prediction pass (backend.cpp:437-600, non-video) for one slice of nbx x nby blocks,
// restated over plain arrays with the reference's own etc_block / color_distance primitives (the member function itself needs a live
// frontend; the drop-in tests compare whole files for that). pIdx: in/out endpoint index per block; pPred: predictor per block,
// 3 = none, 0x83 = none and zero current error.
REF_API void ref_backend_endpoint_prediction(const uint8_t* pBlocks, const uint8_t* pEtc, uint32_t nbx, uint32_t nby, const uint8_t* pC5i, float thresh, uint32_t perceptual,
	uint32_t* pIdx, uint8_t* pPred)
{
	ref_init();
	static const int dx[3] = { -1, 0, -1 }, dy[3] = { 0, -1, -1 };
	for (uint32_t by = 0; by < nby; by++)
		for (uint32_t bx = 0; bx < nbx; bx++)
		{
			const uint32_t b = bx + by * nbx;
			const uint32_t block_endpoint = pIdx[b];
			uint32_t best_pred = UINT32_MAX;
			for (uint32_t p = 0; p < 3; p++)
			{
				const int px = (int)bx + dx[p], py = (int)by + dy[p];
				if (px < 0 || py < 0) continue;
				if (pIdx[px + py * nbx] == block_endpoint && p < best_pred) best_pred = p;
			}
			pPred[b] = 3;
			if (best_pred != UINT32_MAX) { pPred[b] = (uint8_t)best_pred; continue; }
			if (!(thresh > 0.0f)) continue;
			const color_rgba* src = reinterpret_cast<const color_rgba*>(pBlocks + (size_t)b * 64);
			etc_block etc_blk(*reinterpret_cast<const etc_block*>(pEtc + (size_t)b * 8));
			const uint64_t cur_err = etc_blk.evaluate_etc1_error(src, perceptual != 0);
			if (!cur_err) { pPred[b] = 0x83; continue; }
			const uint64_t thresh_err = (uint64_t)(cur_err * maximum(1.0f, thresh));
			etc_block trial(etc_blk);
			uint64_t best_err = UINT64_MAX;
			uint32_t best_index = 0;
			for (uint32_t p = 0; p < 3; p++)
			{
				const int px = (int)bx + dx[p], py = (int)by + dy[p];
				if (px < 0 || py < 0) continue;
				const uint32_t pi = pIdx[px + py * nbx];
				const color_rgba pc(pC5i[pi * 4], pC5i[pi * 4 + 1], pC5i[pi * 4 + 2], 255);
				trial.set_block_color5(pc, pc);
				trial.set_inten_table(0, pC5i[pi * 4 + 3]);
				trial.set_inten_table(1, pC5i[pi * 4 + 3]);
				color_rgba tc[16];
				unpack_etc1(trial, tc);
				uint64_t trial_err = 0;
				for (uint32_t i = 0; i < 16; i++) trial_err += color_distance(perceptual != 0, src[i], tc[i], false);
				if (trial_err <= thresh_err && (trial_err < best_err || (trial_err == best_err && p < best_pred))) { best_pred = p; best_err = trial_err; best_index = pi; }
			}
			if (best_pred != UINT32_MAX) { pIdx[b] = best_index; pPred[b] = (uint8_t)best_pred; }
		}
}

// ---- the steps either side of the per-block path -----------------------------------------------

// basis_compressor::extract_source_blocks for one slice (comp.cpp:3207): image::extract_block_clamped per 4x4 block.
REF_API void ref_extract_source_blocks(const uint8_t* pRGBA, uint32_t w, uint32_t h, uint8_t* pBlocks)
{
	image img(pRGBA, w, h, 4);
	const uint32_t nbx = (w + 3) / 4, nby = (h + 3) / 4;
	for (uint32_t by = 0; by < nby; by++)
		for (uint32_t bx = 0; bx < nbx; bx++)
			img.extract_block_clamped(reinterpret_cast<color_rgba*>(pBlocks + ((size_t)by * nbx + bx) * 64), bx * 4, by * 4, 4, 4);
}

// image_metrics::calc (enc.cpp:2155) on two RGBA images; out = { max, mean, mean_squared, rms, psnr }.
REF_API void ref_image_metrics(const uint8_t* pA, const uint8_t* pB, uint32_t w, uint32_t h, uint32_t first_chan, uint32_t total_chans,
	uint32_t avg_comp_error, uint32_t use_601_luma, double* out5)
{
	image a(pA, w, h, 4), b(pB, w, h, 4);
	image_metrics im;
	im.calc(a, b, first_chan, total_chans, avg_comp_error != 0, use_601_luma != 0);
	out5[0] = im.m_max; out5[1] = im.m_mean; out5[2] = im.m_mean_squared; out5[3] = im.m_rms; out5[4] = im.m_psnr;
}

// The per-block part of basisu_frontend::generate_selector_clusters (frontend.cpp:2156-2179): the block's selector training
// vector (16 selectors, here packed 2 bits each, texel x + 4y at bits 2(x+4y)) and its weight.
REF_API void ref_selector_training(const uint8_t* pEtc_blocks, uint32_t n, uint32_t perceptual, uint32_t* pKeys, uint32_t* pWeights)
{
	for (uint32_t i = 0; i < n; i++)
	{
		const etc_block& blk = *reinterpret_cast<const etc_block*>(pEtc_blocks + (size_t)i * 8);
		uint32_t key = 0;
		for (uint32_t y = 0; y < 4; y++)
			for (uint32_t x = 0; x < 4; x++)
				key |= blk.get_selector(x, y) << ((x + y * 4) * 2);
		const uint32_t subblock_index = (blk.get_inten_table(0) > blk.get_inten_table(1)) ? 0 : 1;
		color_rgba block_colors[2];
		blk.get_block_low_high_colors(block_colors, subblock_index);
		const uint32_t dist = color_distance(perceptual != 0, block_colors[0], block_colors[1], false);
		pKeys[i] = key;
		pWeights[i] = clamp<uint32_t>(dist / 300, 1, 4096);
	}
}

// Turns the encoder's own debug_printf stage timers on or off (enc.h: enable_debug_printf); used by tools/bench_dropin.py.
REF_API void ref_enable_debug_printf(int enabled) { enable_debug_printf(enabled != 0); }

// ---- the ETC1S codebook stages (endpoint/selector VQ, selector codebook) ---------------------------------------

// generate_hierarchical_codebook_threaded (enc.h:2219) over tree_vector_quant<vec6F> (dim 6) or <vec16F> (dim 16), exactly as the
// frontend calls it (frontend.cpp:868, 2140). Outputs are CSR lists of training-vector indices; *_off arrays need n + 1 entries.
template<typename Vec> static int ref_tsvq_t(uint32_t n, const float* pVecs, const uint64_t* pWeights, uint32_t max_codebook, uint32_t max_parent, uint32_t max_threads, bool even_odd,
	uint32_t* cl_off, uint32_t* cl_idx, uint32_t* pNum_clusters, uint32_t* pa_off, uint32_t* pa_idx, uint32_t* pNum_parents)
{
	tree_vector_quant<Vec> q;
	for (uint32_t i = 0; i < n; i++)
	{
		Vec v;
		for (uint32_t c = 0; c < Vec::num_elements; c++) v[c] = pVecs[(size_t)i * Vec::num_elements + c];
		q.add_training_vec(v, pWeights[i]);
	}
	job_pool pool(max_threads ? max_threads : 1);
	basisu::vector<uint_vec> codebook, parent_codebook;
	if (!generate_hierarchical_codebook_threaded(q, max_codebook, max_parent, codebook, parent_codebook, max_threads, &pool, even_odd))
		return 0;
	uint32_t k = 0;
	cl_off[0] = 0;
	for (uint32_t i = 0; i < codebook.size(); i++) { for (uint32_t j = 0; j < codebook[i].size(); j++) cl_idx[k++] = codebook[i][j]; cl_off[i + 1] = k; }
	*pNum_clusters = (uint32_t)codebook.size();
	k = 0;
	pa_off[0] = 0;
	for (uint32_t i = 0; i < parent_codebook.size(); i++) { for (uint32_t j = 0; j < parent_codebook[i].size(); j++) pa_idx[k++] = parent_codebook[i][j]; pa_off[i + 1] = k; }
	*pNum_parents = (uint32_t)parent_codebook.size();
	return 1;
}

REF_API int ref_tsvq(uint32_t dim, uint32_t n, const float* pVecs, const uint64_t* pWeights, uint32_t max_codebook, uint32_t max_parent, uint32_t max_threads, uint32_t even_odd,
	uint32_t* cl_off, uint32_t* cl_idx, uint32_t* pNum_clusters, uint32_t* pa_off, uint32_t* pa_idx, uint32_t* pNum_parents)
{
	ref_init();
	if (dim == 6) return ref_tsvq_t<vec6F>(n, pVecs, pWeights, max_codebook, max_parent, max_threads, even_odd != 0, cl_off, cl_idx, pNum_clusters, pa_off, pa_idx, pNum_parents);
	if (dim == 16) return ref_tsvq_t<vec16F>(n, pVecs, pWeights, max_codebook, max_parent, max_threads, even_odd != 0, cl_off, cl_idx, pNum_clusters, pa_off, pa_idx, pNum_parents);
	return 0;
}

// basisu_frontend::create_optimized_selector_codebook (frontend.cpp:2259-2345) for clusters given as CSR block lists:
// out[c] = the 16 optimised selectors of cluster c, texel (x, y) at bits 2 * (x + 4 * y); 0 for an empty cluster.
REF_API void ref_optimize_selector_codebook(const uint8_t* pPixel_blocks, const uint8_t* pEtc_blocks, uint32_t total_clusters, const uint32_t* pOffsets, const uint32_t* pBlock_indices,
	uint32_t perceptual, uint32_t* pOut)
{
	for (uint32_t c = 0; c < total_clusters; c++)
	{
		uint64_t total_err[4][4][4];
		memset(total_err, 0, sizeof(total_err));
		for (uint32_t k = pOffsets[c]; k < pOffsets[c + 1]; k++)
		{
			const uint32_t bi = pBlock_indices[k];
			const etc_block& blk = *reinterpret_cast<const etc_block*>(pEtc_blocks + (size_t)bi * 8);
			color_rgba blk_colors[4];
			blk.get_block_colors(blk_colors, 0);
			const color_rgba* px = reinterpret_cast<const color_rgba*>(pPixel_blocks + (size_t)bi * 64);
			for (uint32_t y = 0; y < 4; y++)
				for (uint32_t x = 0; x < 4; x++)
					for (uint32_t s = 0; s < 4; s++)
						total_err[y][x][s] += color_distance(perceptual != 0, blk_colors[s], px[x + y * 4], false);
		}
		uint32_t packed = 0;
		if (pOffsets[c + 1] > pOffsets[c])
			for (uint32_t y = 0; y < 4; y++)
				for (uint32_t x = 0; x < 4; x++)
				{
					uint64_t best_err = total_err[y][x][0];
					uint32_t best_sel = 0;
					for (uint32_t s = 1; s < 4; s++)
						if (total_err[y][x][s] < best_err) { best_err = total_err[y][x][s]; best_sel = s; }
					packed |= best_sel << ((x + y * 4) * 2);
				}
		pOut[c] = packed;
	}
}

// ---- mip generation ---------------------------------------------------------------------------------

// basisu::image_resample (enc.cpp:1022) on RGBA8 rasters; pDst holds the destination's initial contents (channels outside the range keep them).
REF_API int ref_image_resample(const uint8_t* pSrc, uint32_t sw, uint32_t sh, uint8_t* pDst, uint32_t dw, uint32_t dh, uint32_t srgb, const char* pFilter, float filter_scale,
	uint32_t wrapping, uint32_t first_comp, uint32_t num_comps)
{
	ref_init();
	image src(pSrc, sw, sh, 4), dst(pDst, dw, dh, 4);
	if (!image_resample(src, dst, srgb != 0, pFilter, filter_scale, wrapping != 0, first_comp, num_comps)) return 0;
	memcpy(pDst, dst.get_ptr(), (size_t)dw * dh * 4);
	return 1;
}

// The contributor lists image_resample's Resampler builds for these sizes (resampler.h:32-42), flattened: offsets[n + 1] + (weight, pixel) pairs.
// Call with pX / pY == NULL to get the counts first. Also the two sRGB tables of enc.cpp:1061-1075.
REF_API int ref_resampler_clists(uint32_t sw, uint32_t sh, uint32_t dw, uint32_t dh, const char* pFilter, float filter_scale, uint32_t wrapping,
	uint32_t* pX_offsets, float* pX_weights, uint32_t* pX_pixels, uint32_t* pY_offsets, float* pY_weights, uint32_t* pY_pixels)
{
	Resampler r(sw, sh, dw, dh, wrapping ? Resampler::BOUNDARY_WRAP : Resampler::BOUNDARY_CLAMP, 0.0f, 1.0f, pFilter, nullptr, nullptr, filter_scale, filter_scale, 0, 0);
	const Resampler::Contrib_List* cx = r.get_clist_x();
	const Resampler::Contrib_List* cy = r.get_clist_y();
	if (!cx || !cy) return 0;
	uint32_t n = 0;
	for (uint32_t i = 0; i < dw; i++)
	{
		pX_offsets[i] = n;
		for (uint32_t j = 0; j < cx[i].n; j++, n++)
			if (pX_weights) { pX_weights[n] = cx[i].p[j].weight; pX_pixels[n] = cx[i].p[j].pixel; }
	}
	pX_offsets[dw] = n;
	n = 0;
	for (uint32_t i = 0; i < dh; i++)
	{
		pY_offsets[i] = n;
		for (uint32_t j = 0; j < cy[i].n; j++, n++)
			if (pY_weights) { pY_weights[n] = cy[i].p[j].weight; pY_pixels[n] = cy[i].p[j].pixel; }
	}
	pY_offsets[dh] = n;
	return 1;
}

REF_API void ref_srgb_tables(float* pSrgb_to_linear256, uint8_t* pLinear_to_srgb8192)
{
	for (int i = 0; i < 256; ++i) pSrgb_to_linear256[i] = srgb_to_linear((float)i * (1.0f / 255.0f));
	for (int i = 0; i < 8192; ++i) pLinear_to_srgb8192[i] = (uint8_t)clamp<int>((int)(255.0f * linear_to_srgb((float)i * (1.0f / (8192 - 1))) + .5f), 0, 255);
}

// palette_index_reorderer::init + get_remap_table (enc.cpp:1785), as the ETC1S backend uses it for the endpoint palette (no distance function).
REF_API void ref_palette_reorder(uint32_t num_indices, const uint32_t* pIndices, uint32_t num_syms, uint32_t* pRemap)
{
	palette_index_reorderer r;
	r.init(num_indices, pIndices, num_syms, nullptr, nullptr, 0);
	const uint_vec& t = r.get_remap_table();
	for (uint32_t i = 0; i < num_syms; i++) pRemap[i] = t[i];
}
