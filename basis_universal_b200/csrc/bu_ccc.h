// bu_ccc.h -- endpoint/selector optimiser for one colour cell (a subset of <= 16 texels), ASTC endpoint ranges,
// linear metric, unit channel weights: the configuration every UASTC mode calls it in.
//
// Behavioural contract: bit-identical results to the reference's
//   color_cell_compression(255, ...)        encoder/basisu_bc7enc.cpp:1364
//     find_optimal_solution                 bc7enc.cpp:1103 (ASTC branch 1107-1163)
//     evaluate_solution                     bc7enc.cpp:822  (non-perceptual branches 899-974)
//     compute_least_squares_endpoints_rgb/a bc7enc.cpp:460 / 394 (double accumulation)
//     pack_astc_*_to_one_color              bc7enc.cpp:628-820
//   color_cell_compression_est_astc         bc7enc.cpp:1764
// including float/double operation order (no FMA contraction), first-strictly-less acceptance and tie rules
// (SURVEY.md 9.5).  The code is organised around packed 32-bit texels and a small state struct rather than
// the reference's params/results structs; BC7-native (non-ASTC) branches do not exist here because UASTC never takes them.
#pragma once
#include "bu_common.h"
#include "bu_tables.h"

// The four table rows the colour-cell optimiser reads are a function of the mode alone (endpoint range, weight bits), and every
// thread of a k_candidates CTA runs the same mode. With BU_STAGE_TABLES the kernel bulk-copies those rows (1056 B) into shared
// memory once per CTA (TMA, cp.async.bulk + mbarrier) and the optimiser reads them there; otherwise they are read from the
// bu_tables aggregate in global memory through L1. Same values either way; tools/ab_bench.py measures the difference.
#if defined(BU_STAGE_TABLES) && defined(BU_UASTC_ENCODER_TU) && defined(__CUDA_ARCH__) // only the translation unit whose kernel stages the rows
namespace bu_stage
{
	struct rows { uint8_t su[256]; uint8_t nearest[256]; uint8_t wt[32]; float wx[128]; };
	__shared__ __align__(128) rows s_rows;
}
#define BU_ROW_SU(T, cfg) (bu_stage::s_rows.su)
#define BU_ROW_NEAREST(T, cfg) (bu_stage::s_rows.nearest)
#define BU_ROW_WT(T, cfg) (bu_stage::s_rows.wt)
#define BU_ROW_WX(T, cfg) (bu_stage::s_rows.wx)
#else
#define BU_ROW_SU(T, cfg) ((T)->sorted_unq + (cfg).slot * 256)
#define BU_ROW_NEAREST(T, cfg) ((T)->nearest + (cfg).slot * 256)
#define BU_ROW_WT(T, cfg) ((T)->weights + (cfg).wbits * 32)
#define BU_ROW_WX(T, cfg) ((T)->weightsx + (cfg).wbits * 32 * 4)
#endif

// The interpolated-colour table of cell_evaluate (<= 32 words, indexed by a data-dependent selector, hence not register-resident):
// thread-local memory by default; with BU_WC_SMEM a per-thread column of a shared-memory array (word i of thread t at [i][t]:
// conflict-free whatever the selectors are), which cannot be evicted to L2 / DRAM the way local-memory lines are.
#if defined(BU_WC_SMEM) && defined(BU_UASTC_ENCODER_TU) && defined(__CUDA_ARCH__)
namespace bu_stage { __shared__ uint32_t s_wc[32 * 128]; } // k_candidates runs 128 threads per CTA
#define BU_WC_DECL uint32_t* const wc_ = bu_stage::s_wc + threadIdx.x;
#define BU_WC(i) wc_[(i) * 128]
#else
#define BU_WC_DECL uint32_t wc_[32];
#define BU_WC(i) wc_[i]
#endif

namespace bu {

struct cell_cfg
{
	uint32_t wbits;      // selector bits: N = 1 << wbits interpolation levels
	uint32_t slot;       // bu_tables range slot of the ASTC endpoint range
	uint32_t range;      // the ASTC endpoint range itself (7,8,11,12,13,18,19,20)
	uint32_t has_alpha;  // 4-component fit if nonzero
	uint32_t uber;       // bc7enc "uber level" (selector perturbation passes)
	uint32_t ls_passes;  // least-squares refinement passes
};

struct cell_fit
{
	uint64_t err;        // best total squared error so far (UINT64_MAX = none)
	uint8_t lo[4], hi[4]; // best endpoints, in sorted-order index space
	uint8_t sel[16];     // best selectors
};

// ---- evaluate one endpoint pair (bc7enc.cpp:822) -------------------------------------------------------------------

BU_NI inline void cell_evaluate(const bu_tables* T, const cell_cfg& cfg, const uint32_t* px, uint32_t n, const uint8_t* lo, const uint8_t* hi, cell_fit& fit, const uint8_t* force_sel = nullptr)
{
	const uint8_t* su = BU_ROW_SU(T, cfg);
	const uint32_t N = 1u << cfg.wbits;
	const uint8_t* wt = BU_ROW_WT(T, cfg);

	int lc[4], dc[4];
	for (int c = 0; c < 4; c++)
	{
		lc[c] = su[lo[c]];
		dc[c] = (int)su[hi[c]] - lc[c];
	}

	// Interpolated colours, channels (0,2) and (1,3) on 16-bit lanes. RGB cells keep byte 3 of the inner entries zero
	// (the RGB metric never reads it).
	BU_WC_DECL
	BU_WC(0) = px_make(lc[0], lc[1], lc[2], lc[3]);
	BU_WC(N - 1) = px_make(lc[0] + dc[0], lc[1] + dc[1], lc[2] + dc[2], lc[3] + dc[3]);
	{
		const uint32_t l02 = (uint32_t)lc[0] | ((uint32_t)lc[2] << 16), h02 = (uint32_t)(lc[0] + dc[0]) | ((uint32_t)(lc[2] + dc[2]) << 16);
		const uint32_t l13 = cfg.has_alpha ? ((uint32_t)lc[1] | ((uint32_t)lc[3] << 16)) : (uint32_t)lc[1];
		const uint32_t h13 = cfg.has_alpha ? ((uint32_t)(lc[1] + dc[1]) | ((uint32_t)(lc[3] + dc[3]) << 16)) : (uint32_t)(lc[1] + dc[1]);
		BU_ROLL
		for (uint32_t i = 1; i + 1 < N; i++)
		{
			const uint32_t w = wt[i];
			BU_WC(i) = astc_lerp_x2(l02, h02, w) | (astc_lerp_x2(l13, h13, w) << 8);
		}
	}

	uint8_t st[16];
	uint32_t total = 0; // <= 16 * 4 * 255^2, fits 32 bits
	const uint32_t cmask = cfg.has_alpha ? 0xFFFFFFFFu : 0x00FFFFFFu;
	if (force_sel)
	{
		// caller-imposed selectors (bc7enc.cpp:885-898): only the error is computed
		BU_ROLL
		for (uint32_t i = 0; i < n; i++)
		{
			const uint32_t s = force_sel[i];
			total += dist_masked(BU_WC(s), px[i], cmask);
			st[i] = (uint8_t)s;
		}
	}
	else
	{
		// Projection of each texel on the endpoint axis picks two neighbouring selectors to compare (bc7enc.cpp:906-972).
		// dot = sum_c (p_c - l_c) d_c = sum_c p_c d_c - K, with K constant per endpoint pair: exact integer arithmetic either way.
		const bool a4 = cfg.has_alpha != 0;
		const int d3 = a4 ? dc[3] : 0;
		const float f = a4 ? (float)N / ((float)(dc[0] * dc[0] + dc[1] * dc[1] + dc[2] * dc[2] + dc[3] * dc[3]) + .00000125f)
		                   : (float)N / ((float)(dc[0] * dc[0] + dc[1] * dc[1] + dc[2] * dc[2]) + .00000125f);
		const uint32_t dc01 = pack_s16x2(dc[0], dc[1]), dc23 = pack_s16x2(dc[2], d3);
		const int K = lc[0] * dc[0] + lc[1] * dc[1] + lc[2] * dc[2] + (a4 ? lc[3] * dc[3] : 0);
		BU_ROLL
		for (uint32_t i = 0; i < n; i++)
		{
			const uint32_t p = px[i];
			const int dot = dot_s16x4_u8x4(dc01, dc23, p) - K;
			int s = (int)((float)dot * f + .5f);
			s = clampi(s, 1, (int)N - 1);
			const uint32_t e0 = dist_masked(BU_WC(s - 1), p, cmask), e1 = dist_masked(BU_WC(s), p, cmask);
			uint32_t e = e1;
			if (e0 == e1) { if (s == 1) s = 0; }   // prefer the non-interpolated endpoint
			else if (e0 < e1) { e = e0; --s; }
			total += e;
			st[i] = (uint8_t)s;
		}
	}

	if (total < fit.err)
	{
		fit.err = total;
		for (int c = 0; c < 4; c++) { fit.lo[c] = lo[c]; fit.hi[c] = hi[c]; }
		BU_ROLL
		for (uint32_t i = 0; i < n; i++) fit.sel[i] = st[i];
	}
}

// ---- quantise float endpoints and try them (bc7enc.cpp:1103, ASTC branch) ------------------------------------------------

BU_FI bool cell_same_endpoints(const cell_fit& fit, const uint8_t* lo, const uint8_t* hi)
{
	for (int c = 0; c < 4; c++)
		if (lo[c] != fit.lo[c] || hi[c] != fit.hi[c]) return false;
	return true;
}

BU_NI inline uint64_t cell_try_endpoints(const bu_tables* T, const cell_cfg& cfg, const uint32_t* px, uint32_t n, const float* xl_in, const float* xh_in, cell_fit& fit, const uint8_t* force_sel = nullptr)
{
	float xl[4], xh[4];
	uint8_t lo0[4], hi0[4];
	const uint8_t* nearest = BU_ROW_NEAREST(T, cfg);
	for (int c = 0; c < 4; c++)
	{
		xl[c] = saturatef_(xl_in[c]);
		xh[c] = saturatef_(xh_in[c]);
		lo0[c] = nearest[clampi((int)(xl[c] * 255.0f + .5f), 0, 255)];
		hi0[c] = nearest[clampi((int)(xh[c] * 255.0f + .5f), 0, 255)];
	}

	bool degenerate = false;
	for (int c = 0; c < 3; c++)
		if (lo0[c] == hi0[c] && fabsf(xl[c] - xh[c]) > 0.0f) degenerate = true;

	if (degenerate)
	{
		const int top = (int)T->range_levels[cfg.slot] - 1;
		// The reference tries nudging low down (1), nothing (0), high up (2), both (3), in that order (bc7enc.cpp:1128-1148).
		for (int k = 0; k < 4; k++)
		{
			const int flags = (k == 0) ? 1 : (k == 1) ? 0 : k;
			uint8_t lo[4], hi[4];
			for (int c = 0; c < 4; c++) { lo[c] = lo0[c]; hi[c] = hi0[c]; }
			for (int c = 0; c < 3; c++)
				if (lo[c] == hi[c] && fabsf(xl[c] - xh[c]) > 0.000125f)
				{
					if ((flags & 1) && lo[c] > 0) lo[c]--;
					if ((flags & 2) && (int)hi[c] < top) hi[c]++;
				}
			if (fit.err == UINT64_MAX || !cell_same_endpoints(fit, lo, hi))
				cell_evaluate(T, cfg, px, n, lo, hi, fit, force_sel);
		}
	}
	else if (fit.err == UINT64_MAX || !cell_same_endpoints(fit, lo0, hi0))
		cell_evaluate(T, cfg, px, n, lo0, hi0, fit, force_sel);

	return fit.err;
}

// ---- least squares endpoints for fixed selectors (bc7enc.cpp:394/460) ----------------------------------------------------

BU_NI inline void cell_least_squares(const bu_tables* T, const cell_cfg& cfg, const uint32_t* px, uint32_t n, const uint8_t* sel, float* xl, float* xh)
{
	const float* wx = BU_ROW_WX(T, cfg);
	const int nc = cfg.has_alpha ? 4 : 3;
	double z00 = 0.0, z10 = 0.0, z11 = 0.0;
	double q00[4] = { 0, 0, 0, 0 }, t[4] = { 0, 0, 0, 0 };

	BU_ROLL
	for (uint32_t i = 0; i < n; i++)
	{
		const float* w4 = wx + sel[i] * 4;
		z00 += w4[0];
		z10 += w4[1];
		z11 += w4[2];
		const float w = w4[3];
		const uint32_t p = px[i];
		for (int c = 0; c < nc; c++)
		{
			const int v = (int)px_c(p, c);
			q00[c] += w * (float)v; // float product, then widened: matches "q00_r += w * pColors[i].m_c[0]"
			t[c] += v;
		}
	}

	const double z01 = z10;
	double det = z00 * z11 - z01 * z10;
	if (det != 0.0) det = 1.0 / det;
	const double iz00 = z11 * det, iz01 = -z01 * det, iz10 = -z10 * det, iz11 = z00 * det;

	for (int c = 0; c < nc; c++)
	{
		const double q10 = t[c] - q00[c];
		xl[c] = (float)(iz00 * q00[c] + iz01 * q10);
		xh[c] = (float)(iz10 * q00[c] + iz11 * q10);
	}
	if (!cfg.has_alpha) { xl[3] = 255.0f; xh[3] = 255.0f; }

	for (int c = 0; c < nc; c++)
		if (xl[c] < 0.0f || xh[c] > 255.0f)
		{
			uint32_t lo_v = 0xFFFFFFFFu, hi_v = 0;
			BU_ROLL
			for (uint32_t i = 0; i < n; i++) { lo_v = minu(lo_v, px_c(px[i], c)); hi_v = maxu(hi_v, px_c(px[i], c)); }
			if (lo_v == hi_v) { xl[c] = (float)lo_v; xh[c] = (float)hi_v; }
		}
}

// LS refit + quantise + evaluate; returns false if a zero-error solution was reached (reference returns early then).
BU_NI inline bool cell_refit(const bu_tables* T, const cell_cfg& cfg, const uint32_t* px, uint32_t n, const uint8_t* sel, cell_fit& fit, const uint8_t* force_sel = nullptr)
{
	float xl[4], xh[4];
	cell_least_squares(T, cfg, px, n, sel, xl, xh);
	for (int c = 0; c < 4; c++) { xl[c] = xl[c] * (1.0f / 255.0f); xh[c] = xh[c] * (1.0f / 255.0f); }
	return cell_try_endpoints(T, cfg, px, n, xl, xh, fit, force_sel) != 0;
}

// ---- single-colour tables (bc7enc.cpp:628-820) ----------------------------------------------------------------------

// Returns the table for this (range, weight bits, alpha) combination or nullptr, plus the selector the table was built for
// and the sorted-order alpha endpoint the reference stores for RGB-only cells.
BU_FI const uint8_t* cell_one_colour_table(const bu_tables* T, const cell_cfg& cfg, uint32_t& sel, uint32_t& alpha_index)
{
	alpha_index = 0;
	if (cfg.range == 8 && cfg.wbits == 3 && !cfg.has_alpha) { sel = 2; return T->one_r8_w3; }
	if (cfg.range == 7 && cfg.wbits == 2 && !cfg.has_alpha) { sel = 1; return T->one_r7_w2; }
	if (cfg.range == 8 && cfg.wbits == 2 && cfg.has_alpha) { sel = 1; return T->one_r8_w2; }
	if (cfg.range == 13 && cfg.wbits == 2 && !cfg.has_alpha) { sel = 1; alpha_index = 47; return T->one_r13_w2; }
	if (cfg.range == 11 && cfg.wbits == 5 && !cfg.has_alpha) { sel = 13; alpha_index = 31; return T->one_r11_w5; }
	return nullptr;
}

BU_NI inline uint64_t cell_one_colour(const bu_tables* T, const cell_cfg& cfg, const uint8_t* tab, uint32_t sel, uint32_t alpha_index,
	const uint32_t* px, uint32_t n, const uint32_t* c4, uint8_t* lo, uint8_t* hi)
{
	const uint8_t* su = BU_ROW_SU(T, cfg);
	const uint32_t w = BU_ROW_WT(T, cfg)[sel];
	for (int c = 0; c < 3; c++) { lo[c] = tab[c4[c] * 2]; hi[c] = tab[c4[c] * 2 + 1]; }
	if (cfg.has_alpha) { lo[3] = tab[c4[3] * 2]; hi[3] = tab[c4[3] * 2 + 1]; }
	else { lo[3] = (uint8_t)alpha_index; hi[3] = (uint8_t)alpha_index; }

	uint32_t p = 0;
	for (int c = 0; c < 3; c++) p |= astc_lerp(su[lo[c]], su[hi[c]], w) << (c * 8);
	// Reference quirks kept: RGB-only range-8/range-7 cells compare with alpha 255, range 13/11 interpolate the alpha slot (unused by the RGB metric).
	if (cfg.has_alpha) p |= astc_lerp(su[lo[3]], su[hi[3]], w) << 24;

	uint32_t total = 0; // <= 16 * 4 * 255^2
	const uint32_t cmask = cfg.has_alpha ? 0xFFFFFFFFu : 0x00FFFFFFu;
	BU_ROLL
	for (uint32_t i = 0; i < n; i++) total += dist_masked(p, px[i], cmask);
	return total;
}

// ---- the full optimiser (bc7enc.cpp:1364) -----------------------------------------------------------------------------

struct cell_result
{
	uint64_t err;
	uint8_t astc_lo[4], astc_hi[4]; // endpoints as ASTC (BISE) indices
	uint8_t sel[16];
};

BU_HD inline void cell_finish(const bu_tables* T, const cell_cfg& cfg, const cell_fit& fit, uint32_t n, cell_result& out)
{
	const uint8_t* si = T->sorted_idx + cfg.slot * 256;
	out.err = fit.err;
	for (int c = 0; c < 4; c++) { out.astc_lo[c] = si[fit.lo[c]]; out.astc_hi[c] = si[fit.hi[c]]; }
	BU_ROLL
	for (uint32_t i = 0; i < n; i++) out.sel[i] = fit.sel[i];
}

// The general path of cell_compress: PCA start, least-squares passes, selector perturbation, mean-colour fallback.
// Kept as its own function with a single exit: the caller's solid-colour shortcut is a data-dependent (divergent) branch, and
// every thread of a warp leaves through the same return.
BU_NI inline void cell_compress_general(const bu_tables* T, const cell_cfg& cfg, const uint32_t* px, uint32_t n, cell_fit& fit,
	const uint8_t* one_tab, uint32_t one_sel, uint32_t one_alpha, const uint8_t* force_sel)
{
	// Mean and principal axis (bc7enc.cpp:1413-1498).
	float mean[4] = { 0, 0, 0, 0 };
	BU_ROLL
	for (uint32_t i = 0; i < n; i++)
		for (int c = 0; c < 4; c++) mean[c] = mean[c] + (float)px_c(px[i], c);

	float mean_s[4], mean_n[4];
	{
		const float inv_n = 1.0f / (float)n;
		const float inv_n255 = 1.0f / ((float)n * 255.0f);
		for (int c = 0; c < 4; c++) { mean_s[c] = mean[c] * inv_n; mean_n[c] = saturatef_(mean[c] * inv_n255); }
	}

	float axis[4];
	if (cfg.has_alpha)
	{
		axis[0] = axis[1] = axis[2] = axis[3] = 0.0f;
		BU_ROLL
		for (uint32_t i = 0; i < n; i++)
		{
			float d[4];
			for (int c = 0; c < 4; c++) d[c] = (float)px_c(px[i], c) - mean_s[c];
			float nv[4];
			for (int c = 0; c < 4; c++) nv[c] = i ? axis[c] : d[c];
			float s = nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2] + nv[3] * nv[3];
			if (s != 0.0f) { s = 1.0f / sqrtf(s); nv[0] *= s; nv[1] *= s; nv[2] *= s; nv[3] *= s; }
			for (int c = 0; c < 4; c++)
			{
				const float k = d[c];
				axis[c] += (d[0] * k) * nv[0] + (d[1] * k) * nv[1] + (d[2] * k) * nv[2] + (d[3] * k) * nv[3];
			}
		}
		float s = axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2] + axis[3] * axis[3];
		if (s != 0.0f) { s = 1.0f / sqrtf(s); axis[0] *= s; axis[1] *= s; axis[2] *= s; axis[3] *= s; }
	}
	else
	{
		float cov[6] = { 0, 0, 0, 0, 0, 0 };
		BU_ROLL
		for (uint32_t i = 0; i < n; i++)
		{
			const float r = (float)px_c(px[i], 0) - mean_s[0], g = (float)px_c(px[i], 1) - mean_s[1], b = (float)px_c(px[i], 2) - mean_s[2];
			cov[0] += r * r; cov[1] += r * g; cov[2] += r * b; cov[3] += g * g; cov[4] += g * b; cov[5] += b * b;
		}
		float xr = .9f, xg = 1.0f, xb = .7f;
		for (int iter = 0; iter < 3; iter++)
		{
			float r = xr * cov[0] + xg * cov[1] + xb * cov[2];
			float g = xr * cov[1] + xg * cov[3] + xb * cov[4];
			float b = xr * cov[2] + xg * cov[4] + xb * cov[5];
			float m = maxf_(maxf_(fabsf(r), fabsf(g)), fabsf(b));
			if (m > 1e-10f) { m = 1.0f / m; r *= m; g *= m; b *= m; }
			xr = r; xg = g; xb = b;
		}
		float len = xr * xr + xg * xg + xb * xb;
		if (len < 1e-10f) { axis[0] = axis[1] = axis[2] = axis[3] = 0.0f; }
		else
		{
			len = 1.0f / sqrtf(len);
			xr *= len; xg *= len; xb *= len;
			axis[0] = xr; axis[1] = xg; axis[2] = xb; axis[3] = 0.0f;
		}
	}

	if (axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2] + axis[3] * axis[3] < .5f)
	{
		axis[0] = axis[1] = axis[2] = 1.0f;
		axis[3] = cfg.has_alpha ? 1.0f : 0.0f;
		float s = axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2] + axis[3] * axis[3];
		if (s != 0.0f) { s = 1.0f / sqrtf(s); axis[0] *= s; axis[1] *= s; axis[2] *= s; axis[3] *= s; }
	}

	float l = 1e+9f, h = -1e+9f;
	BU_ROLL
	for (uint32_t i = 0; i < n; i++)
	{
		float q[4];
		for (int c = 0; c < 4; c++) q[c] = (float)px_c(px[i], c) - mean_s[c];
		const float d = q[0] * axis[0] + q[1] * axis[1] + q[2] * axis[2] + q[3] * axis[3];
		l = minf_(l, d);
		h = maxf_(h, d);
	}
	l *= (1.0f / 255.0f);
	h *= (1.0f / 255.0f);

	float minc[4], maxc[4];
	for (int c = 0; c < 4; c++)
	{
		minc[c] = saturatef_(mean_n[c] + axis[c] * l);
		maxc[c] = saturatef_(mean_n[c] + axis[c] * h);
	}
	// dot with (1,1,1,1): x*1 is exact, so the sums below equal the reference's vec4F_dot results.
	if (minc[0] + minc[1] + minc[2] + minc[3] > maxc[0] + maxc[1] + maxc[2] + maxc[3])
		for (int c = 0; c < 4; c++) { const float tmp = minc[c]; minc[c] = maxc[c]; maxc[c] = tmp; }

	// PCA solution, then least-squares passes (bc7enc.cpp:1546-1565). A zero-error solution ends the search.
	bool go = cell_try_endpoints(T, cfg, px, n, minc, maxc, fit, force_sel) != 0;

	for (uint32_t pass = 0; go && pass < cfg.ls_passes; pass++)
		go = cell_refit(T, cfg, px, n, fit.sel, fit, force_sel);

	if (go && cfg.uber > 0 && !force_sel)
	{
		// Selector perturbation (bc7enc.cpp:1567-1677): bump the extreme selectors inward, refit.
		uint8_t base[16] = { 0 }, trial[16] = { 0 };
		uint32_t min_sel = 256, max_sel = 0;
		const uint32_t top = (1u << cfg.wbits) - 1;
		BU_ROLL
		for (uint32_t i = 0; i < n; i++) { base[i] = fit.sel[i]; min_sel = minu(min_sel, base[i]); max_sel = maxu(max_sel, base[i]); }

		BU_ROLL
		for (uint32_t i = 0; i < n; i++) { uint32_t s = base[i]; if (s == min_sel && s < top) s++; trial[i] = (uint8_t)s; }
		go = cell_refit(T, cfg, px, n, trial, fit);

		if (go)
		{
			BU_ROLL
			for (uint32_t i = 0; i < n; i++) { uint32_t s = base[i]; if (s == max_sel && s > 0) s--; trial[i] = (uint8_t)s; }
			go = cell_refit(T, cfg, px, n, trial, fit);
		}
		if (go)
		{
			BU_ROLL
			for (uint32_t i = 0; i < n; i++)
			{
				uint32_t s = base[i];
				if (s == min_sel && s < top) s++;
				else if (s == max_sel && s > 0) s--;
				trial[i] = (uint8_t)s;
			}
			go = cell_refit(T, cfg, px, n, trial, fit);
		}

		// Uber >= 2: rescale the selector range to exploit endpoint extrapolation (bc7enc.cpp:1647-1677).
		const uint32_t thresh = (n * 56) >> 4;
		if (go && cfg.uber >= 2 && fit.err > thresh)
		{
			const int Q = (cfg.uber >= 4) ? ((int)cfg.uber - 2) : 1;
			const int max_selector = (int)top;
			for (int ly = -Q; go && ly <= 1; ly++)
				for (int hy = max_selector - 1; go && hy <= max_selector + Q; hy++)
				{
					if (ly == 0 && hy == max_selector) continue;
					BU_ROLL
					for (uint32_t i = 0; i < n; i++)
						trial[i] = f2u8_x86(clampf_(floorf((float)max_selector * ((float)base[i] - (float)ly) / ((float)hy - (float)ly) + .5f), 0.0f, (float)max_selector));
					go = cell_refit(T, cfg, px, n, trial, fit);
				}
		}
	}

	// Finally try coding the cell's mean as a single colour (bc7enc.cpp:1680-1755). Skipped when an earlier step hit zero error
	// (the reference has already returned by then).
	if (go && one_tab)
	{
		uint32_t c4[4];
		for (int c = 0; c < 4; c++) c4[c] = (uint32_t)(int)(.5f + mean_n[c] * 255.0f);
		uint8_t lo[4], hi[4];
		const uint64_t avg_err = cell_one_colour(T, cfg, one_tab, one_sel, one_alpha, px, n, c4, lo, hi);
		if (avg_err < fit.err)
		{
			fit.err = avg_err;
			for (int c = 0; c < 4; c++) { fit.lo[c] = lo[c]; fit.hi[c] = hi[c]; }
			BU_ROLL
			for (uint32_t i = 0; i < n; i++) fit.sel[i] = (uint8_t)one_sel;
		}
	}

}

// force_sel != nullptr: the selectors are imposed (uastc_rdo endpoint refinement): no single-colour shortcuts, no selector perturbation.
BU_NI inline void cell_compress(const bu_tables* T, const cell_cfg& cfg, const uint32_t* px, uint32_t n, cell_result& out, const uint8_t* force_sel = nullptr)
{
	cell_fit fit;
	fit.err = UINT64_MAX;
	for (int c = 0; c < 4; c++) { fit.lo[c] = 0; fit.hi[c] = 0; }
	BU_ROLL
	for (int i = 0; i < 16; i++) fit.sel[i] = 0;

	uint32_t one_sel = 0, one_alpha = 0;
	const uint8_t* one_tab = force_sel ? nullptr : cell_one_colour_table(T, cfg, one_sel, one_alpha);

	// All texels equal and a single-colour table exists for this configuration: done (bc7enc.cpp:1379-1411).
	bool solid = false;
	if (one_tab)
	{
		const uint32_t mask = cfg.has_alpha ? 0xFFFFFFFFu : 0x00FFFFFFu;
		uint32_t differing = 0;
		BU_ROLL
		for (uint32_t i = 1; i < n; i++) differing |= (px[i] ^ px[0]) & mask;
		solid = differing == 0;
	}
	if (solid)
	{
		const uint32_t c4[4] = { px_c(px[0], 0), px_c(px[0], 1), px_c(px[0], 2), px_c(px[0], 3) };
		fit.err = cell_one_colour(T, cfg, one_tab, one_sel, one_alpha, px, n, c4, fit.lo, fit.hi);
		BU_ROLL
		for (uint32_t i = 0; i < n; i++) fit.sel[i] = (uint8_t)one_sel;
	}
	else
		cell_compress_general(T, cfg, px, n, fit, one_tab, one_sel, one_alpha, force_sel);

	cell_finish(T, cfg, fit, n, out);
}

// ---- fast bounding-box estimate used to rank partitions (bc7enc.cpp:1764) --------------------------------------------------
// Integer only. The reference's early-out on best_err_so_far only truncates sums that already exceed the running best, so
// returning the full sum leaves every comparison made by the partition rankers unchanged.

#if defined(BU_RANK_INLINE)
BU_FI uint64_t cell_estimate(
#else
BU_NI inline uint64_t cell_estimate(
#endif
const bu_tables* T, uint32_t wbits, uint32_t comps, const uint32_t* px, uint32_t n)
{
	const uint32_t N = 1u << wbits; // wbits is 2 or 3 here
	const uint8_t* wt = T->weights + wbits * 32;
	uint32_t lo4 = 0xFFFFFFFFu, hi4 = 0;
	BU_ROLL
	for (uint32_t i = 0; i < n; i++) { lo4 = min_u8x4(lo4, px[i]); hi4 = max_u8x4(hi4, px[i]); }
	if (comps == 3) { lo4 |= 0xFF000000u; hi4 |= 0xFF000000u; }
	const uint32_t lo[4] = { px_c(lo4, 0), px_c(lo4, 1), px_c(lo4, 2), px_c(lo4, 3) }, hi[4] = { px_c(hi4, 0), px_c(hi4, 1), px_c(hi4, 2), px_c(hi4, 3) };

	uint32_t wc[8];
	wc[0] = lo4;
	wc[N - 1] = hi4;
	{
		const uint32_t l02 = lo[0] | (lo[2] << 16), h02 = hi[0] | (hi[2] << 16);
		const uint32_t l13 = (comps == 4) ? (lo[1] | (lo[3] << 16)) : lo[1], h13 = (comps == 4) ? (hi[1] | (hi[3] << 16)) : hi[1];
		const uint32_t opaque = (comps == 4) ? 0u : 0xFF000000u;
		BU_ROLL
		for (uint32_t i = 1; i + 1 < N; i++)
			wc[i] = astc_lerp_x2(l02, h02, wt[i]) | (astc_lerp_x2(l13, h13, wt[i]) << 8) | opaque;
	}

	// Projection axis = bounding-box diagonal; selector thresholds = midpoints of the interpolated colours' projections.
	// Only 2- and 3-bit weights are ranked (N <= 8); unused thresholds stay at INT_MAX.
	const uint32_t a01 = pack_s16x2((int)hi[0] - (int)lo[0], (int)hi[1] - (int)lo[1]);
	const uint32_t a23 = pack_s16x2((int)hi[2] - (int)lo[2], (comps == 4) ? (int)hi[3] - (int)lo[3] : 0);
	int thresh[7] = { 0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF };
	{
		int prev = 0;
		BU_ROLL
		for (uint32_t i = 0; i < N; i++)
		{
			const int dot = dot_s16x4_u8x4(a01, a23, wc[i]);
			if (i) thresh[i - 1] = (prev + dot + 1) >> 1;
			prev = dot;
		}
	}

	uint64_t total = 0;
	const uint32_t cmask = (comps == 4) ? 0xFFFFFFFFu : 0x00FFFFFFu;
	BU_ROLL
	for (uint32_t i = 0; i < n; i++)
	{
		const uint32_t p = px[i];
		const int d = dot_s16x4_u8x4(a01, a23, p);
		// The reference scans the thresholds from the top for the first one <= d. They are non-decreasing (bounding-box
		// endpoints: every channel of wc[] is non-decreasing in the weight, and a[] >= 0), so that index is a count.
		uint32_t s = 0;
		if (N == 4) s = ((d >= thresh[0]) ? 1u : 0u) + ((d >= thresh[1]) ? 1u : 0u) + ((d >= thresh[2]) ? 1u : 0u);
		else for (uint32_t j = 0; j < 7; j++) s += (d >= thresh[j]) ? 1u : 0u;
		total += dist_masked(wc[s], p, cmask);
	}
	return total;
}

} // namespace bu
