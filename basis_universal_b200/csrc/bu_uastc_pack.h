// bu_uastc_pack.h -- transcode hints (BC1, ETC2 EAC A8, ETC1) for the chosen candidate and the final 128-bit packing.
//
// Bit-exact with, respectively: compute_bc1_hints (encoder/basisu_uastc_enc.cpp:2535) built on basist::encode_bc1
// (transcoder/basisu_transcoder.cpp:18047), transcode_uastc_to_bc1_hint0/1 (18602/18700) and bcu::unpack_bc1
// (basisu_dds_transcoder.inl:23); uastc_pack_eac_a8 (uastc_enc.cpp:3019); compute_etc1_hints (2714) with
// pack_etc1_estimate_flipped (2668) and apply_etc1_bias (transcoder.cpp:16547); pack_etc1_block_solid_color
// (encoder/basisu_etc.cpp:181); pack_uastc (uastc_enc.cpp:110).
//
// The ETC1 hint search is exposed per (flip, individual/differential) *trial group* so the GPU path can run the four groups
// of a block on four threads and reduce with an (error, trial order) arg-min, which equals the reference's sequential
// first-strictly-less rule.
#pragma once
#include "bu_uastc.h"

namespace bu {

// ---- canonical (anchor-normalised) form of a candidate: what pack_uastc writes and unpack_uastc reads back -------------------
// uastc_enc.cpp:267-339: per plane and subset, if the anchor texel's weight has its MSB set, invert that subset's weights in
// that plane and swap the endpoints that plane controls.

BU_NI inline void canonicalize(const bu_tables* T, const candidate& c, uint8_t* ep, uint8_t* w)
{
	const uint32_t mode = c.mode;
	const uint32_t subsets = T->mode_subsets[mode], comps = T->mode_comps[mode], planes = T->mode_planes[mode];
	const uint32_t wbits = T->mode_weight_bits[mode];
	BU_ROLL
	for (int i = 0; i < 18; i++) ep[i] = c.ep[i];
	BU_ROLL
	for (uint32_t i = 0; i < 16 * planes; i++) w[i] = c.w[i];

	const uint8_t zero16[16] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
	const uint8_t zero3[3] = { 0, 0, 0 };
	const uint8_t* part = (subsets >= 2) ? astc_partition(T, mode, c.pattern) : zero16;
	const uint8_t* anchors = (subsets >= 2) ? astc_anchors(T, mode, c.pattern) : zero3;

	BU_ROLL
	for (uint32_t plane = 0; plane < planes; plane++)
		BU_ROLL
		for (uint32_t s = 0; s < subsets; s++)
		{
			const uint32_t anchor = anchors[s];
			if (!(w[anchor * planes + plane] & (1u << (wbits - 1)))) continue;
			BU_ROLL
			for (int i = 0; i < 16; i++)
				if (part[i] == s) w[i * planes + plane] = (uint8_t)(((1u << wbits) - 1) - w[i * planes + plane]);
			if (planes == 2)
			{
				for (uint32_t k = 0; k < comps; k++)
				{
					const uint32_t comp_plane = (comps == 2) ? k : ((k == c.ccs) ? 1u : 0u);
					if (comp_plane == plane) { const uint8_t t = ep[k * 2]; ep[k * 2] = ep[k * 2 + 1]; ep[k * 2 + 1] = t; }
				}
			}
			else
				for (uint32_t k = 0; k < comps; k++)
				{
					uint8_t* e = ep + s * comps * 2 + k * 2;
					const uint8_t t = e[0]; e[0] = e[1]; e[1] = t;
				}
		}
}

// ---- BC1 ----------------------------------------------------------------------------------------------------------------

struct bc1_logical { uint32_t c0, c1; uint32_t sel; }; // 565 colours (c0 = "low"/first field) and 16 2-bit raw selectors

BU_NI inline void bc1_decode(const bc1_logical& b, uint32_t* out)
{
	const uint32_t l = b.c0, h = b.c1;
	uint32_t r0 = (l >> 11) & 31, g0 = (l >> 5) & 63, b0 = l & 31, r1 = (h >> 11) & 31, g1 = (h >> 5) & 63, b1 = h & 31;
	r0 = (r0 << 3) | (r0 >> 2); g0 = (g0 << 2) | (g0 >> 4); b0 = (b0 << 3) | (b0 >> 2);
	r1 = (r1 << 3) | (r1 >> 2); g1 = (g1 << 2) | (g1 >> 4); b1 = (b1 << 3) | (b1 >> 2);
	uint32_t c[4];
	c[0] = px_make(r0, g0, b0, 255);
	c[1] = px_make(r1, g1, b1, 255);
	if (l > h)
	{
		c[2] = px_make((r0 * 2 + r1) / 3, (g0 * 2 + g1) / 3, (b0 * 2 + b1) / 3, 255);
		c[3] = px_make((r1 * 2 + r0) / 3, (g1 * 2 + g0) / 3, (b1 * 2 + b0) / 3, 255);
	}
	else
	{
		c[2] = px_make((r0 + r1) / 2, (g0 + g1) / 2, (b0 + b1) / 2, 255);
		c[3] = 0;
	}
	BU_ROLL
	for (int i = 0; i < 16; i++) out[i] = c[(b.sel >> (i * 2)) & 3];
}

// transcoder.cpp:17857
BU_NI inline void bc1_find_sels(const uint32_t* px, uint32_t lr, uint32_t lg, uint32_t lb, uint32_t hr, uint32_t hg, uint32_t hb, uint8_t* sels)
{
	uint32_t br[4], bg[4], bb[4];
	br[0] = (lr << 3) | (lr >> 2); bg[0] = (lg << 2) | (lg >> 4); bb[0] = (lb << 3) | (lb >> 2);
	br[3] = (hr << 3) | (hr >> 2); bg[3] = (hg << 2) | (hg >> 4); bb[3] = (hb << 3) | (hb >> 2);
	br[1] = (br[0] * 2 + br[3]) / 3; bg[1] = (bg[0] * 2 + bg[3]) / 3; bb[1] = (bb[0] * 2 + bb[3]) / 3;
	br[2] = (br[3] * 2 + br[0]) / 3; bg[2] = (bg[3] * 2 + bg[0]) / 3; bb[2] = (bb[3] * 2 + bb[0]) / 3;
	int ar = (int)br[3] - (int)br[0], ag = (int)bg[3] - (int)bg[0], ab = (int)bb[3] - (int)bb[0];
	int dots[4];
	for (int i = 0; i < 4; i++) dots[i] = (int)br[i] * ar + (int)bg[i] * ag + (int)bb[i] * ab;
	const int t0 = dots[0] + dots[1], t1 = dots[1] + dots[2], t2 = dots[2] + dots[3];
	ar *= 2; ag *= 2; ab *= 2;
	BU_ROLL
	for (int i = 0; i < 16; i++)
	{
		const int d = (int)px_c(px[i], 0) * ar + (int)px_c(px[i], 1) * ag + (int)px_c(px[i], 2) * ab;
		sels[i] = (uint8_t)(3 - ((d <= t0) + (d < t1) + (d < t2)));
	}
}

// transcoder.cpp:17922. Returns false when the normal equations are singular.
BU_NI inline bool bc1_least_squares(const uint32_t* px, const uint8_t* sels, float* xl, float* xh)
{
	uint32_t uq00[3] = { 0, 0, 0 }, ut[3] = { 0, 0, 0 };
	const uint32_t wv[4] = { 0x000009, 0x010204, 0x040201, 0x090000 };
	uint32_t wacc = 0;
	BU_ROLL
	for (int i = 0; i < 16; i++)
	{
		const uint32_t s = sels[i];
		wacc += wv[s];
		for (int c = 0; c < 3; c++) { const uint32_t v = px_c(px[i], c); ut[c] += v; uq00[c] += s * v; }
	}
	const float z00 = (float)((wacc >> 16) & 0xFF), z10 = (float)((wacc >> 8) & 0xFF), z11 = (float)(wacc & 0xFF), z01 = z10;
	float det = z00 * z11 - z01 * z10;
	if (fabsf(det) < 1e-8f) return false;
	det = 3.0f / det;
	const float iz00 = z11 * det, iz01 = -z01 * det, iz10 = -z10 * det, iz11 = z00 * det;
	for (int c = 0; c < 3; c++)
	{
		const float q00 = (float)uq00[c], t = (float)ut[c];
		const float q10 = t * 3.0f - q00;
		xl[c] = iz00 * q00 + iz01 * q10;
		xh[c] = iz10 * q00 + iz11 * q10;
	}
	for (int c = 0; c < 3; c++)
		if (xl[c] < 0.0f || xh[c] > 255.0f)
		{
			uint32_t lo_v = 0xFFFFFFFFu, hi_v = 0;
			BU_ROLL
			for (int i = 0; i < 16; i++) { lo_v = minu(lo_v, px_c(px[i], c)); hi_v = maxu(hi_v, px_c(px[i], c)); }
			if (lo_v == hi_v) { xl[c] = (float)lo_v; xh[c] = (float)hi_v; }
		}
	return true;
}

BU_FI uint32_t to_5(uint32_t v) { v = v * 31 + 128; return (v + (v >> 8)) >> 8; }
BU_FI uint32_t to_6(uint32_t v) { v = v * 63 + 128; return (v + (v >> 8)) >> 8; }

// encode_bc1 (transcoder.cpp:18047) with flags = 0 (use_sels == false) or cEncodeBC1UseSelectors (raw selectors supplied).
BU_NI inline void bc1_encode(const bu_tables* T, const uint32_t* px, bool use_sels, uint32_t raw_sels_in, bc1_logical& out)
{
	int avg[3] = { -1, 0, 0 };
	int lr = 0, lg = 0, lb = 0, hr = 0, hg = 0, hb = 0;
	uint8_t sels[16];

	if (use_sels)
	{
		const uint8_t tran[4] = { 0, 3, 1, 2 };
		BU_ROLL
		for (int i = 0; i < 16; i++) sels[i] = tran[(raw_sels_in >> (i * 2)) & 3];
	}
	else
	{
		const uint32_t fr = px_c(px[0], 0), fg = px_c(px[0], 1), fb = px_c(px[0], 2);
		int j;
		BU_ROLL
		for (j = 1; j < 16; j++)
			if (((px[j] ^ px[0]) & 0x00FFFFFFu) != 0) break;
		if (j == 16)
		{
			// encode_bc1_solid_block (transcoder.cpp:17999)
			uint32_t mask = 0xAA;
			uint32_t max16 = ((uint32_t)T->bc1_match5[fr * 2] << 11) | ((uint32_t)T->bc1_match6[fg * 2] << 5) | T->bc1_match5[fb * 2];
			uint32_t min16 = ((uint32_t)T->bc1_match5[fr * 2 + 1] << 11) | ((uint32_t)T->bc1_match6[fg * 2 + 1] << 5) | T->bc1_match5[fb * 2 + 1];
			if (min16 == max16)
			{
				mask = 0;
				if (min16 > 0) min16--;
				else { max16 = 1; min16 = 0; mask = 0x55; }
			}
			if (max16 < min16) { const uint32_t t = max16; max16 = min16; min16 = t; mask ^= 0x55; }
			out.c0 = max16; out.c1 = min16; out.sel = mask * 0x01010101u;
			return;
		}

		int total[3] = { (int)fr, (int)fg, (int)fb }, mx[3] = { (int)fr, (int)fg, (int)fb }, mn[3] = { (int)fr, (int)fg, (int)fb };
		BU_ROLL
		for (int i = 1; i < 16; i++)
			for (int c = 0; c < 3; c++)
			{
				const int v = (int)px_c(px[i], c);
				mx[c] = maxi(mx[c], v); mn[c] = mini(mn[c], v); total[c] += v;
			}
		for (int c = 0; c < 3; c++) avg[c] = (total[c] + 8) >> 4;

		int icov[6] = { 0, 0, 0, 0, 0, 0 };
		BU_ROLL
		for (int i = 0; i < 16; i++)
		{
			const int r = (int)px_c(px[i], 0) - avg[0], g = (int)px_c(px[i], 1) - avg[1], b = (int)px_c(px[i], 2) - avg[2];
			icov[0] += r * r; icov[1] += r * g; icov[2] += r * b; icov[3] += g * g; icov[4] += g * b; icov[5] += b * b;
		}
		float cov[6];
		for (int i = 0; i < 6; i++) cov[i] = (float)icov[i] * (1.0f / 255.0f);

		float xr = (float)(mx[0] - mn[0]), xg = (float)(mx[1] - mn[1]), xb = (float)(mx[2] - mn[2]);
		for (int it = 0; it < 4; it++)
		{
			const float r = xr * cov[0] + xg * cov[1] + xb * cov[2];
			const float g = xr * cov[1] + xg * cov[3] + xb * cov[4];
			const float b = xr * cov[2] + xg * cov[4] + xb * cov[5];
			xr = r; xg = g; xb = b;
		}
		const float k = maxf_(maxf_(fabsf(xr), fabsf(xg)), fabsf(xb));
		int sr = 306, sg = 601, sb = 117;
		if (k >= 2.0f)
		{
			const float m = 1024.0f / k;
			sr = (int)(xr * m); sg = (int)(xg * m); sb = (int)(xb * m);
		}
		int low_dot = 2147483647, high_dot = (-2147483647 - 1), low_c = 0, high_c = 0;
		BU_ROLL
		for (int i = 0; i < 16; i++)
		{
			const int dot = (int)px_c(px[i], 0) * sr + (int)px_c(px[i], 1) * sg + (int)px_c(px[i], 2) * sb;
			if (dot < low_dot) { low_dot = dot; low_c = i; }
			if (dot > high_dot) { high_dot = dot; high_c = i; }
		}
		lr = (int)to_5(px_c(px[low_c], 0)); lg = (int)to_6(px_c(px[low_c], 1)); lb = (int)to_5(px_c(px[low_c], 2));
		hr = (int)to_5(px_c(px[high_c], 0)); hg = (int)to_6(px_c(px[high_c], 1)); hb = (int)to_5(px_c(px[high_c], 2));
		bc1_find_sels(px, lr, lg, lb, hr, hg, hb, sels);
	}

	// one least-squares pass (flags never carry the high-quality bits on this path)
	{
		float xl[3], xh[3];
		if (!bc1_least_squares(px, sels, xl, xh))
		{
			if (avg[0] < 0)
			{
				int total[3] = { 0, 0, 0 };
				for (int i = 0; i < 16; i++) for (int c = 0; c < 3; c++) total[c] += (int)px_c(px[i], c);
				for (int c = 0; c < 3; c++) avg[c] = (total[c] + 8) >> 4;
			}
			lr = T->bc1_match5[avg[0] * 2]; lg = T->bc1_match6[avg[1] * 2]; lb = T->bc1_match5[avg[2] * 2];
			hr = T->bc1_match5[avg[0] * 2 + 1]; hg = T->bc1_match6[avg[1] * 2 + 1]; hb = T->bc1_match5[avg[2] * 2 + 1];
		}
		else
		{
			lr = clampi((int)(xl[0] * (31.0f / 255.0f) + .5f), 0, 31);
			lg = clampi((int)(xl[1] * (63.0f / 255.0f) + .5f), 0, 63);
			lb = clampi((int)(xl[2] * (31.0f / 255.0f) + .5f), 0, 31);
			hr = clampi((int)(xh[0] * (31.0f / 255.0f) + .5f), 0, 31);
			hg = clampi((int)(xh[1] * (63.0f / 255.0f) + .5f), 0, 63);
			hb = clampi((int)(xh[2] * (31.0f / 255.0f) + .5f), 0, 31);
		}
		bc1_find_sels(px, lr, lg, lb, hr, hg, hb, sels);
	}

	uint32_t lc16 = (uint32_t)lb | ((uint32_t)lg << 5) | ((uint32_t)lr << 11);
	uint32_t hc16 = (uint32_t)hb | ((uint32_t)hg << 5) | ((uint32_t)hr << 11);
	if (lc16 == hc16)
	{
		uint32_t mask = 0;
		if (hc16 > 0) hc16--;
		else { hc16 = 0; lc16 = 1; mask = 0x55; }
		out.c0 = lc16; out.c1 = hc16; out.sel = mask * 0x01010101u;
	}
	else
	{
		uint32_t inv = 0;
		if (lc16 < hc16) { const uint32_t t = lc16; lc16 = hc16; hc16 = t; inv = 0x55555555u; }
		const uint8_t tr[4] = { 0, 2, 3, 1 };
		uint32_t packed = 0;
		BU_ROLL
		for (int i = 0; i < 16; i++) packed |= (uint32_t)tr[sels[i]] << (i * 2);
		out.c0 = lc16; out.c1 = hc16; out.sel = packed ^ inv;
	}
}

BU_FI uint32_t pack565_scaled(uint32_t r, uint32_t g, uint32_t b)
{
	r = (r * 31u + 127u) / 255u; g = (g * 63u + 127u) / 255u; b = (b * 31u + 127u) / 255u;
	return minu(b, 31u) | (minu(g, 63u) << 5) | (minu(r, 31u) << 11);
}

// compute_bc1_hints (uastc_enc.cpp:2535). ep/w are the canonical endpoints/weights of the chosen candidate.
BU_NI inline void compute_bc1_hints(const bu_tables* T, uint32_t mode, const uint8_t* ep, const uint8_t* w, const uint32_t* src, const uint32_t* dec, bool& hint0, bool& hint1)
{
	hint0 = false; hint1 = false;
	const bool has0 = T->mode_has_bc1_hint0[mode] != 0, has1 = T->mode_has_bc1_hint1[mode] != 0;
	if (!has0 && !has1) return;

	uint32_t t_px[16], h0_px[16], h1_px[16];
	bc1_logical b;
	bc1_encode(T, dec, false, 0, b);
	bc1_decode(b, t_px);

	const uint32_t wbits = T->mode_weight_bits[mode], plane_shift = T->mode_planes[mode] - 1;
	const uint8_t* tran = T->uastc_to_bc1 + wbits * 32;

	if (has1)
	{
		// transcode_uastc_to_bc1_hint1: first-plane weights mapped to BC1 selectors, endpoints by least squares
		uint32_t sels = 0;
		BU_ROLL
		for (int i = 15; i >= 0; --i) sels = (sels << 2) | tran[w[i << plane_shift]];
		bc1_encode(T, dec, true, sels, b);
		bc1_decode(b, h1_px);
	}
	else for (int i = 0; i < 16; i++) h1_px[i] = 0;

	if (has0)
	{
		// transcode_uastc_to_bc1_hint0: first-subset endpoints scaled to 565, weights mapped directly
		const uint8_t* u = T->unq + (uint32_t)T->range_slot[T->mode_endpoint_range[mode]] * 256;
		uint32_t c0, c1;
		if (T->mode_comps[mode] == 2)
		{
			const uint32_t l = u[ep[0]], h = u[ep[1]];
			c0 = pack565_scaled(l, l, l); c1 = pack565_scaled(h, h, h);
		}
		else
		{
			c0 = pack565_scaled(u[ep[0]], u[ep[2]], u[ep[4]]);
			c1 = pack565_scaled(u[ep[1]], u[ep[3]], u[ep[5]]);
		}
		if (c0 == c1)
		{
			uint32_t mask = 0;
			if (c1 > 0) c1--;
			else { c1 = 0; c0 = 1; mask = 0x55; }
			b.c0 = c0; b.c1 = c1; b.sel = mask * 0x01010101u;
		}
		else
		{
			bool invert = false;
			if (c0 < c1) { const uint32_t t = c0; c0 = c1; c1 = t; invert = true; }
			uint32_t sels = 0;
			BU_ROLL
			for (int i = 15; i >= 0; --i)
			{
				uint32_t s = tran[w[i << plane_shift]];
				if (invert) s ^= 1;
				sels = (sels << 2) | s;
			}
			b.c0 = c0; b.c1 = c1; b.sel = sels;
		}
		bc1_decode(b, h0_px);
	}
	else for (int i = 0; i < 16; i++) h0_px[i] = 0;

	uint64_t et = 0, e0 = 0, e1 = 0;
	BU_ROLL
	for (int i = 0; i < 16; i++) { et += dist_rgb(src[i], t_px[i]); e0 += dist_rgb(src[i], h0_px[i]); e1 += dist_rgb(src[i], h1_px[i]); }
	const float ft = sqrtf((float)et), f0 = sqrtf((float)e0), f1 = sqrtf((float)e1);
	if (has0 && f0 <= ft * 1.075f) hint0 = true;
	if (has1 && f1 <= ft * 1.075f) hint1 = true;
}

// ---- ETC2 EAC A8 hint (uastc_enc.cpp:3019, base_search_rad = 0) ------------------------------------------------------------

BU_NI inline void compute_eac_hint(const bu_tables* T, const uint32_t* dec, uint32_t mul_rad, uint32_t table_mask, uint32_t& out_table, uint32_t& out_mul)
{
	uint32_t mn = 255, mx = 0;
	BU_ROLL
	for (int i = 0; i < 16; i++) { const uint32_t a = px_c(dec[i], 3); if (a < mn) mn = a; if (a > mx) mx = a; }
	if (mn == mx) { out_table = 13; out_mul = 1; return; }
	const uint32_t alpha_range = mx - mn;
	uint64_t best = UINT64_MAX;
	out_table = 0; out_mul = 0;
	BU_ROLL
	for (uint32_t table = 0; table < 16; table++)
	{
		if (!(table_mask & (1u << table))) continue;
		const int8_t* tab = T->eac_tables + table * 8;
		const float range = (float)((int)tab[7] - (int)tab[3]);
		const float fmn = (float)mn, fmx = (float)mx;
		const int center = (int)roundf(fmn + (fmx - fmn) * ((float)(0 - (int)tab[3]) / range));
		const int base = clamp255i(center);
		const int mul = (int)roundf((float)alpha_range / range);
		const int mul_low = clampi(mul - (int)mul_rad, 1, 15), mul_high = clampi(mul + (int)mul_rad, 1, 15);
		BU_ROLL
		for (int m = mul_low; m <= mul_high; m++)
		{
			uint64_t total = 0;
			BU_ROLL
			for (int i = 0; i < 16; i++)
			{
				const int a = (int)px_c(dec[i], 3);
				uint32_t be = 0xFFFFFFFFu;
				for (int s = 0; s < 8; s++)
				{
					const uint32_t e = (uint32_t)iabsi(a - clamp255i(m * (int)tab[s] + base));
					if (e < be) be = e;
				}
				total += (uint64_t)(be * be);
			}
			if (total < best)
			{
				best = total; out_mul = (uint32_t)m; out_table = table;
				if (!best) return;
			}
		}
	}
}

// ---- ETC1 hints (uastc_enc.cpp:2714) -------------------------------------------------------------------------------------

struct ycc { int y, cb, cr; };
BU_FI ycc to_ycc(uint32_t p)
{
	ycc v;
	const int r = (int)px_c(p, 0), g = (int)px_c(p, 1), b = (int)px_c(p, 2);
	v.y = r * 54 + g * 183 + b * 19;
	v.cb = (b << 8) - v.y;
	v.cr = (r << 8) - v.y;
	return v;
}
BU_FI uint64_t sq_u64(int d) { const uint32_t a = (uint32_t)(d < 0 ? -d : d); return (uint64_t)a * (uint64_t)a; } // one IMAD.WIDE.U32
// uastc_enc.cpp:2645 color_diff: 4*dy^2 + dcr^2 + dcb^2 in 64 bits (|d| < 2^18, so each square needs up to 35 bits).
// Device form: (2 dy)^2 + dcr^2 + dcb^2 as three chained signed 32x32->64 multiply-adds.
BU_FI uint64_t ycc_diff(const ycc& a, const ycc& b)
{
#if defined(__CUDA_ARCH__)
	const int dy2 = (a.y - b.y) * 2, dcr = a.cr - b.cr, dcb = a.cb - b.cb;
	uint64_t r;
	asm("{\n\t.reg .s64 t;\n\tmul.wide.s32 t, %1, %1;\n\tmad.wide.s32 t, %2, %2, t;\n\tmad.wide.s32 t, %3, %3, t;\n\tmov.b64 %0, t;\n\t}" : "=l"(r) : "r"(dy2), "r"(dcr), "r"(dcb));
	return r;
#else
	return sq_u64(a.y - b.y) * 4 + sq_u64(a.cr - b.cr) + sq_u64(a.cb - b.cb);
#endif
}

BU_FI int gray_distance2(uint32_t p, int r, int g, int b)
{
	const int c0 = (int)px_c(p, 0), c1 = (int)px_c(p, 1), c2 = (int)px_c(p, 2);
	const int gd = ((c0 - r) + (c1 - g) + (c2 - b) + 1) / 3;
	const int d0 = c0 - clamp255i(r + gd), d1 = c1 - clamp255i(g + gd), d2 = c2 - clamp255i(b + gd);
	return d0 * d0 + d1 * d1 + d2 * d2;
}

// uastc_enc.cpp:2668
BU_NI inline bool etc1_estimate_flipped(const uint32_t* px)
{
	int upper[3], lower[3], left[3], right[3];
	for (uint32_t c = 0; c < 3; c++)
	{
		const int s00 = (int)(px_c(px[0], c) + px_c(px[4], c) + px_c(px[1], c) + px_c(px[5], c));
		const int s10 = (int)(px_c(px[2], c) + px_c(px[6], c) + px_c(px[3], c) + px_c(px[7], c));
		const int s01 = (int)(px_c(px[8], c) + px_c(px[12], c) + px_c(px[9], c) + px_c(px[13], c));
		const int s11 = (int)(px_c(px[10], c) + px_c(px[14], c) + px_c(px[11], c) + px_c(px[15], c));
		upper[c] = (s00 + s10 + 4) / 8; lower[c] = (s01 + s11 + 4) / 8; left[c] = (s00 + s01 + 4) / 8; right[c] = (s10 + s11 + 4) / 8;
	}
	int ud = 0, ld = 0, lfd = 0, rd = 0;
	BU_ROLL
	for (int i = 0; i < 4; i++)
		for (int j = 0; j < 2; j++)
		{
			ud += gray_distance2(px[i + j * 4], upper[0], upper[1], upper[2]);
			ld += gray_distance2(px[i + (2 + j) * 4], lower[0], lower[1], lower[2]);
			lfd += gray_distance2(px[j + i * 4], left[0], left[1], left[2]);
			rd += gray_distance2(px[(2 + j) + i * 4], right[0], right[1], right[2]);
		}
	return (ud + ld) < (lfd + rd);
}

// transcoder.cpp:16547
// The reference's per-bias switch (channel / subblock -> delta in {-2..1}), tabulated: [bias][subblock][channel].
BU_TABLE(int8_t, etc1_bias_deltas, [32][2][3], {
		{ { -1, -1, -1 }, { -1, -1, -1 } }, { {  0, -1, -1 }, {  0, -1, -1 } }, { { -1,  0,  0 }, {  0,  0,  0 } }, { { -1,  0, -1 }, { -1,  0, -1 } },
		{ {  0,  0, -1 }, {  0,  0, -1 } }, { {  0, -1,  0 }, {  0,  0,  0 } }, { {  0,  0, -1 }, {  0,  0,  0 } }, { {  1,  0,  0 }, {  0,  0,  0 } },
		{ {  0,  0,  0 }, {  0,  0,  1 } }, { { -1, -1,  0 }, { -1, -1,  0 } }, { { -2, -2, -2 }, { -2, -2, -2 } }, { {  0,  1,  0 }, {  0,  0,  0 } },
		{ { -1,  0,  0 }, { -1,  0,  0 } }, { {  0,  0,  0 }, {  0,  0,  0 } }, { {  1,  0,  0 }, {  1,  0,  0 } }, { {  0,  0,  1 }, {  0,  0,  0 } },
		{ {  0,  1,  0 }, {  0,  1,  0 } }, { {  1,  1,  0 }, {  1,  1,  0 } }, { {  0,  0,  0 }, { -1,  0,  0 } }, { {  0,  0,  0 }, {  0, -1,  0 } },
		{ {  0,  0,  0 }, {  0,  0, -1 } }, { {  0,  0,  0 }, {  1,  0,  0 } }, { {  0,  0,  1 }, {  0,  0,  1 } }, { {  1,  0,  1 }, {  1,  0,  1 } },
		{ {  0,  0,  0 }, {  0,  1,  0 } }, { {  0,  1,  1 }, {  0,  1,  1 } }, { {  1,  1,  1 }, {  1,  1,  1 } }, { { -1, -1, -1 }, {  0,  0,  0 } },
		{ {  1,  1,  1 }, { -1, -1, -1 } }, { {  0,  0,  0 }, {  1,  1,  1 } }, { {  0,  0,  0 }, { -1, -1, -1 } }, { {  1,  1,  1 }, {  0,  0,  0 } } })
BU_FI int etc1_bias_delta(uint32_t bias, uint32_t subblock, uint32_t c) { return BU_TABLE_REF(etc1_bias_deltas)[bias][subblock][c]; }

// Order in which the bias values are tried at levels 0-3 (uastc_enc.cpp:2730).
BU_TABLE(uint8_t, etc1_sorted_bias, [32], { 13, 0, 22, 29, 27, 12, 26, 9, 30, 31, 8, 10, 25, 2, 23, 5, 15, 7, 3, 11, 6, 17, 28, 18, 1, 19, 20, 21, 24, 4, 14, 16 })

// transcoder.cpp:16547
BU_HD inline void etc1_apply_bias(const int* in, uint32_t bias, int limit, uint32_t subblock, int* out)
{
	for (uint32_t c = 0; c < 3; c++)
	{
		const int delta = etc1_bias_delta(bias, subblock, c);
		int v = in[c];
		if (v == 0) { if (delta == -2) v += 3; else v += delta + 1; }
		else if (v == limit) v += (delta - 1);
		else
		{
			v += delta;
			if (v < 0 || v > limit) v = (v - delta) - delta;
		}
		out[c] = v;
	}
}

struct etc1_hint
{
	uint64_t err;      // UINT64_MAX = no trial accepted
	uint32_t order;    // position of the accepted trial in the reference's (flip, individ, bias_iter) loop nest: the tie-breaker
	uint8_t flip, diff, inten0, inten1, bias;
};

struct etc1_search_opts
{
	uint32_t last_bias;
	uint8_t sorted_bias_table, has_bias, all_inten_tables;
	uint8_t first_flip, last_flip, first_individ, last_individ;
};

// uastc_enc.cpp:2724-2800: trial ranges from level, flags and mode.
BU_HD inline etc1_search_opts etc1_search_setup(const bu_tables* T, uint32_t mode, int level, uint32_t flags, const uint32_t* dec)
{
	etc1_search_opts o;
	const bool faster = (flags & 64) != 0, fastest = (flags & 128) != 0;
	o.has_bias = T->mode_has_etc1_bias[mode];
	o.last_bias = 1; o.sorted_bias_table = 0;
	const bool flip_estimate = (level <= 1) || faster || fastest;
	if (o.has_bias)
	{
		o.sorted_bias_table = 1;
		switch (level)
		{
		case 0: o.last_bias = fastest ? 1 : (faster ? 1 : 2); break;
		case 1: o.last_bias = fastest ? 1 : (faster ? 3 : 5); break;
		case 2: o.last_bias = fastest ? 1 : (faster ? 10 : 20); break;
		case 3: o.last_bias = fastest ? 1 : (faster ? 16 : 32); break;
		default: o.last_bias = 32; o.sorted_bias_table = 0; break;
		}
	}
	o.all_inten_tables = (level == 4);
	o.first_flip = 0; o.last_flip = 2; o.first_individ = 0; o.last_individ = 2;
	if (flags & 256) { o.last_flip = 1; o.last_individ = 1; }
	else if (flip_estimate)
	{
		if (etc1_estimate_flipped(dec)) o.first_flip = 1;
		o.last_flip = o.first_flip + 1;
	}
	return o;
}

// Intensity-table scan and error contribution of one subset for a given scaled base colour (uastc_enc.cpp:2859-2995).
// An ETC1 table colour is base + m on every channel, so unless a channel clamps, its luma is y_base + 256 m and its chroma equals
// the base's: the four candidates for a texel differ in the luma term only and
//     color_diff = 4 (dy + 256 m)^2 + dcr^2 + dcb^2
// is the same integer expression the reference evaluates, factored. A clamping table (base + m outside 0..255) takes the general form.
// Texel of subset s, slot j (g_etc1_pixel_coords, etc.cpp:314): subsets are row pairs when flipped, else column pairs.
BU_FI uint32_t etc1_subset_texel(uint32_t flip, uint32_t s, uint32_t j) { return flip ? (s * 8 + j) : (s * 2 + (j >> 2) + (j & 3) * 4); }

// The loops below are deliberately kept rolled where the trip count is data dependent or the body is large: the stage is
// bound by instruction fetch (ncu: "no instruction" is the top stall), so hot code has to stay within the 32 KB L1.5 I-cache.
BU_NI inline void etc1_subset_search(const bu_tables* T, const etc1_search_opts& o, uint32_t flip, uint32_t s, uint32_t base_key, const int* mn, const int* mx,
	const ycc* src_y, const ycc* dec_y, uint32_t& inten_out, uint64_t& err_out)
{
	const int base[3] = { (int)(base_key & 255), (int)((base_key >> 8) & 255), (int)((base_key >> 16) & 255) };
	int range = 0, bmin = 255, bmax = 0;
	for (int c = 0; c < 3; c++)
	{
		range = maxi(range, iabsi(mx[c] - base[c]));
		range = maxi(range, iabsi(base[c] - mn[c]));
		bmin = mini(bmin, base[c]); bmax = maxi(bmax, base[c]);
	}
	const uint32_t limit = o.all_inten_tables ? 8u : ((range > 51) ? 8u : (range >= 7 ? 4u : 2u));
	const ycc base_y = to_ycc(px_make((uint32_t)base[0], (uint32_t)base[1], (uint32_t)base[2], 255));

	// The four modifiers of a table are (-a, -b, b, a), so min_k (dy + 256 m_k)^2 = (min(||dy| - 256 a|, ||dy| - 256 b|))^2.
	int ady[8];
	uint64_t chroma_sum = 0;
	for (uint32_t j = 0; j < 8; j++)
	{
		const ycc& q = dec_y[etc1_subset_texel(flip, s, j)];
		ady[j] = iabsi(base_y.y - q.y);
		chroma_sum += sq_u64(base_y.cr - q.cr) + sq_u64(base_y.cb - q.cb);
	}

	uint32_t inten = 0;
	uint64_t best_sub = UINT64_MAX;
BU_ROLL
	for (uint32_t t = 0; t < limit; t++)
	{
		const int mb = 256 * (int)T->etc1_inten[t * 4 + 2], ma = 256 * (int)T->etc1_inten[t * 4 + 3]; // small / large positive modifier
		uint64_t total = 0;
		if (bmin * 256 >= ma && bmax * 256 + ma <= 255 * 256)
		{
			for (int j = 0; j < 8; j++)
			{
				const uint32_t d = (uint32_t)mini(iabsi(ady[j] - ma), iabsi(ady[j] - mb));
				total += (uint64_t)d * d;
			}
			total = total * 4 + chroma_sum;
		}
		else
		{
			ycc tab[4];
			for (int k = 0; k < 4; k++)
			{
				const int m = T->etc1_inten[t * 4 + k];
				tab[k] = to_ycc(px_make(clamp255i(base[0] + m), clamp255i(base[1] + m), clamp255i(base[2] + m), 255));
			}
BU_ROLL
			for (uint32_t j = 0; j < 8; j++)
			{
				const ycc q = dec_y[etc1_subset_texel(flip, s, j)];
				total += minu64(minu64(ycc_diff(tab[0], q), ycc_diff(tab[1], q)), minu64(ycc_diff(tab[2], q), ycc_diff(tab[3], q)));
			}
		}
		if (total < best_sub) { best_sub = total; inten = t; }
		// Reference quirk (uastc_enc.cpp:2933): in the non-flipped layout the early-out sits outside the row loop, so the
		// first table that fails to improve ends the whole table scan; in the flipped layout it only ends that table's rows.
		else if (!flip) break;
	}

	// Error of the chosen table against the ORIGINAL block, selectors chosen against the decoded UASTC block (lowest index on ties).
	ycc tab[4];
	bool clamped = false;
	for (int k = 0; k < 4; k++)
	{
		const int m = T->etc1_inten[inten * 4 + k];
		if (bmin + m < 0 || bmax + m > 255) clamped = true;
		tab[k] = to_ycc(px_make(clamp255i(base[0] + m), clamp255i(base[1] + m), clamp255i(base[2] + m), 255));
	}
	uint64_t err = 0;
BU_ROLL
	for (uint32_t j = 0; j < 8; j++)
	{
		const uint32_t i = etc1_subset_texel(flip, s, j);
		const ycc q = dec_y[i];
		ycc chosen;
		if (!clamped)
		{
			// chroma terms are equal for the four candidates: the arg-min is decided by |luma difference| alone, and the chosen
			// colour is base + m on every channel, i.e. (y + 256 m, cr, cb) in the metric's space
			const int dy = base_y.y - q.y;
			int bm = (int)T->etc1_inten[inten * 4];
			int be = iabsi(dy + 256 * bm);
			for (uint32_t k = 1; k < 4; k++) { const int m = (int)T->etc1_inten[inten * 4 + k]; const int e = iabsi(dy + 256 * m); if (e < be) { be = e; bm = m; } }
			chosen.y = base_y.y + 256 * bm; chosen.cr = base_y.cr; chosen.cb = base_y.cb;
		}
		else
		{
			uint32_t bi = 0;
			uint64_t be = ycc_diff(tab[0], q);
			for (uint32_t k = 1; k < 4; k++) { const uint64_t e = ycc_diff(tab[k], q); if (e < be) { be = e; bi = k; } }
			chosen = (bi == 0) ? tab[0] : (bi == 1) ? tab[1] : (bi == 2) ? tab[2] : tab[3];
		}
		err += ycc_diff(src_y[i], chosen);
	}
	inten_out = inten;
	err_out = err;
}

// Trials bias_iter = first, first + stride, ... of one (flip, individ) combination; best kept with first-strictly-less.
// Splitting a combination's bias iterations over threads and reducing by (err, order) reproduces the sequential result.
//
// Both the chosen intensity table and the error contribution of a subset are pure functions of its scaled base colour (for
// fixed flip / texels), and the bias trials revisit the same base colours many times (the 20 sorted biases of level 2 give
// only 13 / 10 distinct per-subset deltas). The work is therefore phased: (1) enumerate the trials' base colours and
// deduplicate them per subset, (2) search each distinct base colour once -- a loop every lane of a warp walks together,
// which a look-aside cache inside the trial loop would not give -- (3) combine per trial in the reference's order.
BU_NI inline void etc1_hint_trials(const bu_tables* T, const etc1_search_opts& o, uint32_t flip, uint32_t individ, uint32_t first, uint32_t stride,
	const ycc* src_y, const ycc* dec_y, const uint32_t* dec, etc1_hint& best)
{
	const int mul = individ ? 15 : 31;

	// subset = row pair when flipped, else column pair (g_etc1_pixel_coords, etc.cpp:314)
	int unbiased[2][3];
	int mn[2][3], mx[2][3];
	BU_ROLL
	for (int s = 0; s < 2; s++)
	{
		uint32_t sum[3] = { 0, 0, 0 };
		for (int c = 0; c < 3; c++) { mn[s][c] = 255; mx[s][c] = 0; }
		for (int j = 0; j < 8; j++)
		{
			const int x = flip ? (j & 3) : (s * 2 + (j >> 2)), y = flip ? (s * 2 + (j >> 2)) : (j & 3);
			const uint32_t p = dec[x + y * 4];
			for (int c = 0; c < 3; c++)
			{
				const int v = (int)px_c(p, c);
				sum[c] += (uint32_t)v;
				mn[s][c] = mini(mn[s][c], v); mx[s][c] = maxi(mx[s][c], v);
			}
		}
		for (int c = 0; c < 3; c++) unbiased[s][c] = (int)(uint8_t)((sum[c] * (uint32_t)mul + 1020) / (8 * 255));
	}

	// (1) base colours of every trial, deduplicated per subset
	uint32_t keys[2][32], n_keys[2] = { 0, 0 };
	uint8_t trial_key[2][32];
	uint32_t n_trials = 0;
	BU_ROLL
	for (uint32_t bias_iter = first; bias_iter < o.last_bias; bias_iter += stride, n_trials++)
	{
		const uint32_t bias = o.sorted_bias_table ? BU_TABLE_REF(etc1_sorted_bias)[bias_iter] : bias_iter;

		int col[2][3];
		for (int s = 0; s < 2; s++)
		{
			if (o.has_bias) etc1_apply_bias(unbiased[s], bias, mul, (uint32_t)s, col[s]);
			else for (int c = 0; c < 3; c++) col[s][c] = unbiased[s][c];
		}

		// Stored base colours -> scaled 8-bit base per subset (etc.h set_block_color4 / set_block_color5_clamp, get_block_color).
		uint32_t key[2] = { 0, 0 };
		if (individ)
		{
			for (int s = 0; s < 2; s++)
				for (int c = 0; c < 3; c++) { const int v = mini(col[s][c], 15); key[s] |= (uint32_t)((v << 4) | v) << (8 * c); }
		}
		else
		{
			for (int c = 0; c < 3; c++)
			{
				const int b0 = mini(col[0][c], 31);
				const int d = clampi(col[1][c] - col[0][c], -4, 3);
				const int b1 = clampi(b0 + d, 0, 31);
				key[0] |= (uint32_t)((b0 << 3) | (b0 >> 2)) << (8 * c);
				key[1] |= (uint32_t)((b1 << 3) | (b1 >> 2)) << (8 * c);
			}
		}
		for (int s = 0; s < 2; s++)
		{
			uint32_t k = 0;
			while (k < n_keys[s] && keys[s][k] != key[s]) k++;
			if (k == n_keys[s]) keys[s][n_keys[s]++] = key[s];
			trial_key[s][n_trials] = (uint8_t)k;
		}
	}

	// (2) one search per distinct base colour
	uint8_t key_inten[2][32];
	uint64_t key_err[2][32];
	for (int s = 0; s < 2; s++)
		BU_ROLL
		for (uint32_t k = 0; k < n_keys[s]; k++)
		{
			uint32_t t; uint64_t e;
			etc1_subset_search(T, o, flip, (uint32_t)s, keys[s][k], mn[s], mx[s], src_y, dec_y, t, e);
			key_inten[s][k] = (uint8_t)t; key_err[s][k] = e;
		}

	// (3) trials in the reference's order, first strictly smaller error wins
	uint32_t bias_iter = first;
	BU_ROLL
	for (uint32_t i = 0; i < n_trials; i++, bias_iter += stride)
	{
		const uint32_t k0 = trial_key[0][i], k1 = trial_key[1][i];
		const uint64_t err = key_err[0][k0] + key_err[1][k1];
		if (err < best.err)
		{
			best.err = err;
			best.order = (flip * 2 + individ) * 32 + bias_iter;
			best.flip = (uint8_t)flip; best.diff = (uint8_t)(individ ? 0 : 1);
			best.inten0 = key_inten[0][k0]; best.inten1 = key_inten[1][k1];
			best.bias = (uint8_t)(o.sorted_bias_table ? BU_TABLE_REF(etc1_sorted_bias)[bias_iter] : bias_iter);
		}
	}
}

// ---- ETC1 solid colour (etc.cpp:181) -> the fields pack_uastc stores for mode 8 --------------------------------------------

struct etc1_solid { uint32_t diff, inten, selector, r, g, b; };

BU_NI inline etc1_solid etc1_pack_solid(const bu_tables* T, uint32_t colour)
{
	const uint32_t next_comp[4] = { 1, 2, 0, 1 };
	uint32_t best_err = 0xFFFFFFFFu, best_i = 0, best_x = 0, best_c1 = 0, best_c2 = 0;
	bool done = false;
	for (uint32_t i = 0; i < 3 && !done; i++)
	{
		const uint32_t c1 = px_c(colour, next_comp[i]), c2 = px_c(colour, next_comp[i + 1]);
		for (int delta = -1; delta <= 1 && !done; delta++)
		{
			const int cpd = clampi((int)px_c(colour, i) + delta, 0, 255);
			const uint16_t* p = T->solid_cfg + T->solid_cfg_ofs[cpd];
			do
			{
				const uint32_t x = *p++;
				const uint16_t* inv = T->etc1_inverse + (x & 0xFF) * 256;
				const uint32_t p1 = inv[c1], p2 = inv[c2];
				const int d = cpd - (int)px_c(colour, i);
				const uint32_t e = (uint32_t)(d * d) + (p1 >> 8) * (p1 >> 8) + (p2 >> 8) * (p2 >> 8);
				if (e < best_err)
				{
					best_err = e; best_x = x; best_c1 = p1 & 0xFF; best_c2 = p2 & 0xFF; best_i = i;
					if (!best_err) { done = true; break; }
				}
			} while (*p != 0xFFFF);
		}
	}
	etc1_solid s;
	s.diff = best_x & 1; s.inten = (best_x >> 1) & 7; s.selector = (best_x >> 4) & 3;
	uint32_t comp[3];
	comp[best_i] = (best_x >> 8) & 255;
	comp[next_comp[best_i]] = best_c1;
	comp[next_comp[best_i + 1]] = best_c2;
	s.r = comp[0]; s.g = comp[1]; s.b = comp[2];
	return s;
}

// ---- bit packing (uastc_enc.cpp:110) ---------------------------------------------------------------------------------------

struct bit_writer
{
	uint64_t lo, hi;
	uint32_t ofs;
};
BU_FI void bw_put(bit_writer& b, uint64_t code, uint32_t n)
{
	if (!n) return;
	if (b.ofs < 64)
	{
		b.lo |= code << b.ofs;
		if (b.ofs + n > 64) b.hi |= code >> (64 - b.ofs);
	}
	else b.hi |= code << (b.ofs - 64);
	b.ofs += n;
}

BU_NI inline void pack_solid_block(const bu_tables* T, uint32_t colour, uint8_t* out16)
{
	bit_writer b; b.lo = 0; b.hi = 0; b.ofs = 0;
	bw_put(b, T->mode_huff[8 * 2], T->mode_huff[8 * 2 + 1]);
	for (uint32_t c = 0; c < 4; c++) bw_put(b, px_c(colour, c), 8);
	const etc1_solid s = etc1_pack_solid(T, colour);
	bw_put(b, s.diff, 1); bw_put(b, s.inten, 3); bw_put(b, s.selector, 2);
	bw_put(b, s.r, 5); bw_put(b, s.g, 5); bw_put(b, s.b, 5);
	for (int i = 0; i < 8; i++) { out16[i] = (uint8_t)(b.lo >> (i * 8)); out16[8 + i] = (uint8_t)(b.hi >> (i * 8)); }
}

// ep/w: canonical endpoints and weights (canonicalize()).
BU_NI inline void pack_block(const bu_tables* T, const candidate& c, const uint8_t* ep, const uint8_t* w, const etc1_hint& etc1,
	uint32_t eac_table, uint32_t eac_mul, bool bc1_hint0, bool bc1_hint1, uint8_t* out16)
{
	const uint32_t mode = c.mode;
	bit_writer b; b.lo = 0; b.hi = 0; b.ofs = 0;
	bw_put(b, T->mode_huff[mode * 2], T->mode_huff[mode * 2 + 1]);
	if (T->mode_has_bc1_hint0[mode]) bw_put(b, bc1_hint0 ? 1 : 0, 1);
	if (T->mode_has_bc1_hint1[mode]) bw_put(b, bc1_hint1 ? 1 : 0, 1);
	bw_put(b, etc1.flip, 1); bw_put(b, etc1.diff, 1); bw_put(b, etc1.inten0, 3); bw_put(b, etc1.inten1, 3);
	if (T->mode_has_etc1_bias[mode]) bw_put(b, etc1.bias, 5);
	if (T->mode_has_alpha[mode]) bw_put(b, eac_table | (eac_mul << 4), 8);

	const uint32_t subsets = T->mode_subsets[mode], planes = T->mode_planes[mode], comps = T->mode_comps[mode];
	if (subsets == 2) bw_put(b, c.pattern, 5);
	else if (subsets == 3) bw_put(b, c.pattern, 4);
	if (mode == 6 || mode == 11 || mode == 13) bw_put(b, c.ccs, 2);

	// endpoints: BISE trit/quint groups first, then the plain bits
	const uint32_t total_values = comps * 2 * subsets;
	const uint32_t range = T->mode_endpoint_range[mode];
	const uint32_t ep_bits = T->bise[range * 3], ep_trits = T->bise[range * 3 + 1], ep_quints = T->bise[range * 3 + 2];
	if (ep_trits || ep_quints)
	{
		const uint32_t radix = ep_trits ? 3u : 5u, full = ep_trits ? 243u : 125u;
		uint32_t accum = 0, mulv = 1;
		BU_ROLL
		for (uint32_t i = 0; i < total_values; i++)
		{
			accum += (uint32_t)(ep[i] >> ep_bits) * mulv;
			mulv *= radix;
			if (mulv == full) { bw_put(b, accum, ep_trits ? 8 : 7); accum = 0; mulv = 1; }
		}
		if (mulv > 1)
		{
			uint32_t nb;
			if (ep_trits) nb = (mulv == 3) ? 2 : (mulv == 9) ? 4 : (mulv == 27) ? 5 : 7;
			else nb = (mulv == 5) ? 3 : 5;
			bw_put(b, accum, nb);
		}
	}
	BU_ROLL
	for (uint32_t i = 0; i < total_values; i++) bw_put(b, ep[i] & ((1u << ep_bits) - 1), ep_bits);

	// weights, anchors one bit short
	const uint32_t wbits = T->mode_weight_bits[mode];
	const uint8_t zero3[3] = { 0, 0, 0 };
	const uint8_t* anchors = (subsets >= 2) ? astc_anchors(T, mode, c.pattern) : zero3;
	const uint32_t plane_shift = planes - 1;
	BU_ROLL
	for (uint32_t i = 0; i < 16 * planes; i++)
	{
		uint32_t nb = wbits;
		for (uint32_t s = 0; s < subsets; s++)
			if (anchors[s] == (i >> plane_shift)) { nb--; break; }
		bw_put(b, w[i], nb);
	}
	for (int i = 0; i < 8; i++) { out16[i] = (uint8_t)(b.lo >> (i * 8)); out16[8 + i] = (uint8_t)(b.hi >> (i * 8)); }
}

} // namespace bu
