// bu_uastc.h -- UASTC LDR 4x4 candidate generation, candidate scoring (UASTC error + BC7-transcode error) and final
// block assembly (BC1 / ETC1 / EAC transcode hints, bit packing).
//
// GPU-first decomposition: the reference's encode_uastc (encoder/basisu_uastc_enc.cpp:3126) is one
// 4000-line scalar routine per block.  Here a block is a set of independent *candidate slots* -- (mode, partition rank or
// plane rotation) pairs -- each generated and scored by its own thread, followed by a per-block select/hint/pack step.
// Every function below is a pure function of its inputs and is bit-exact with the reference routine it cites.
#pragma once
#include "bu_ccc.h"

namespace bu {

// One candidate encoding: UASTC mode + the logical ASTC description (endpoints as BISE indices, weights as indices).
struct candidate
{
	uint8_t mode;       // UASTC mode 0..18
	uint8_t pattern;    // common partition pattern index (2/3-subset modes)
	uint8_t ccs;        // colour component selector (dual-plane modes)
	uint8_t valid;
	uint8_t ep[18];     // RR GG BB [AA] per subset, low/high interleaved
	uint8_t w[32];      // one plane: w[0..15]; two planes: w[2i] plane 0, w[2i+1] plane 1
	uint16_t pad;
	uint32_t uastc_err; // squared error of the UASTC decode vs the source block (metric per block class)
	uint32_t bc7_err;   // squared error of the BC7 transcode of this candidate
};

// Block classification (uastc_enc.cpp:3135-3151, 3271-3277).
struct block_class
{
	uint8_t solid, has_alpha, is_la, pad;
};

BU_HD inline block_class classify_block(const uint32_t* px, bool only_use_la_on_transparent_blocks)
{
	block_class k;
	k.solid = 1; k.has_alpha = 0; k.is_la = 1; k.pad = 0;
	BU_ROLL
	for (int i = 0; i < 16; i++)
	{
		const uint32_t p = px[i];
		if (px_c(p, 3) < 255) k.has_alpha = 1;
		if (p != px[0]) k.solid = 0;
		if (px_c(p, 0) != px_c(p, 1) || px_c(p, 0) != px_c(p, 2)) k.is_la = 0;
	}
	if (only_use_la_on_transparent_blocks && k.is_la && !k.has_alpha) k.is_la = 0;
	return k;
}

// Per-level encoder options (uastc_enc.cpp:3189-3255).
struct level_opts
{
	uint32_t mode_mask, uber, ls_passes;
	uint32_t eac_mul_rad, eac_table_mask;
	uint8_t estimate_partition, always_try_alpha, bc1_hints, la_only_transparent;
};

BU_HD inline level_opts make_level_opts(int level)
{
	level_opts o;
	o.mode_mask = 0xFFFFFFFFu; o.uber = 6; o.ls_passes = 2; o.eac_mul_rad = 3; o.eac_table_mask = 0xFFFFFFFFu;
	o.estimate_partition = 0; o.always_try_alpha = 1; o.bc1_hints = 1; o.la_only_transparent = 0;
	level = clampi(level, 0, 4);
	if (level == 0)
	{
		o.mode_mask = (1u << 0) | (1u << 8) | (1u << 11) | (1u << 12) | (1u << 15);
		o.always_try_alpha = 0; o.eac_mul_rad = 0; o.eac_table_mask = (1u << 2) | (1u << 8) | (1u << 11) | (1u << 13);
		o.uber = 0; o.ls_passes = 1; o.bc1_hints = 0; o.estimate_partition = 1; o.la_only_transparent = 1;
	}
	else if (level == 1)
	{
		o.mode_mask = (1u << 0) | (1u << 4) | (1u << 6) | (1u << 8) | (1u << 9) | (1u << 11) | (1u << 12) | (1u << 15) | (1u << 17);
		o.always_try_alpha = 0; o.eac_mul_rad = 0; o.eac_table_mask = (1u << 2) | (1u << 8) | (1u << 11) | (1u << 13);
		o.uber = 0; o.ls_passes = 1; o.estimate_partition = 1;
	}
	else if (level == 2)
	{
		o.mode_mask = (1u << 0) | (1u << 1) | (1u << 4) | (1u << 5) | (1u << 6) | (1u << 8) | (1u << 9) | (1u << 10) | (1u << 11) | (1u << 12) | (1u << 13) | (1u << 15) | (1u << 16) | (1u << 17);
		o.always_try_alpha = 0; o.eac_mul_rad = 1;
		o.eac_table_mask = (1u << 0) | (1u << 2) | (1u << 6) | (1u << 7) | (1u << 8) | (1u << 10) | (1u << 11) | (1u << 13);
		o.uber = 1; o.ls_passes = 1; o.estimate_partition = 1;
	}
	else if (level == 3)
	{
		o.always_try_alpha = 0; o.eac_mul_rad = 2; o.uber = 3; o.estimate_partition = 1;
	}
	return o;
}

// ---- partition helpers -----------------------------------------------------------------------------------------------

// Texel -> subset in the numbering the colour-cell fits are run in: BC7 numbering for modes 2,3,4,9,16 (uastc_enc.cpp:697,
// 874, 1019, 1670), ASTC numbering for mode 7 (1425: bc7_convert_partition_index_3_to_2 of the BC7 pattern == astc_pat73).
BU_FI const uint8_t* fit_partition(const bu_tables* T, uint32_t mode, uint32_t pattern)
{
	if (mode == 3) return T->bc7_part3 + pattern * 16;
	if (mode == 7) return T->astc_pat73 + pattern * 16;
	return T->bc7_part2 + pattern * 16;
}
// Texel -> subset in ASTC numbering (what the decoder and packer use; transcoder.cpp:15816-15825).
BU_FI const uint8_t* astc_partition(const bu_tables* T, uint32_t mode, uint32_t pattern)
{
	if (mode == 3) return T->astc_pat3 + pattern * 16;
	if (mode == 7) return T->astc_pat73 + pattern * 16;
	return T->astc_pat2 + pattern * 16;
}
BU_FI const uint8_t* astc_anchors(const bu_tables* T, uint32_t mode, uint32_t pattern)
{
	if (mode == 3) return T->anchors3 + pattern * 3;
	if (mode == 7) return T->anchors73 + pattern * 3;
	return T->anchors2 + pattern * 3;
}

// transcoder.cpp:14303
BU_FI uint32_t bc7_part3_to_2(uint32_t p, uint32_t k)
{
	switch (k >> 1)
	{
	case 0: p = (p <= 1) ? 0 : 1; break;
	case 1: p = (p == 0) ? 0 : 1; break;
	default: p = (p == 0 || p == 2) ? 0 : 1; break;
	}
	if (k & 1) p = 1 - p;
	return p;
}

// ---- partition ranking (uastc_enc.cpp:638 estimate_partition2, 1542 estimate_partition2_list, 834-858, 1368-1406) --------

BU_HD inline uint64_t estimate_pattern_error(const bu_tables* T, const uint8_t* part, uint32_t subsets, uint32_t wbits, uint32_t comps, const uint32_t* px)
{
	uint64_t total = 0;
	BU_ROLL
	for (uint32_t s = 0; s < subsets; s++)
	{
		uint32_t sub[16];
		uint32_t n = 0;
		BU_ROLL
		for (int i = 0; i < 16; i++) if (part[i] == s) sub[n++] = px[i];
		total += cell_estimate(T, wbits, comps, sub, n);
	}
	return total;
}

// kind: 2 = two-subset common patterns (30), 3 = three-subset (11), 7 = BC7-3/ASTC-2 patterns (19).
// Writes the best `want` (<= 8) pattern indices in ascending error order, first-strictly-less wins among equals.
BU_NI inline void rank_partitions(const bu_tables* T, uint32_t kind, uint32_t wbits, uint32_t comps, const uint32_t* px, uint32_t want, uint8_t* out)
{
	const uint32_t total = (kind == 2) ? 30u : (kind == 3) ? 11u : 19u;
	const uint8_t* base = (kind == 2) ? T->bc7_part2 : (kind == 3) ? T->bc7_part3 : T->astc_pat73;
	const uint32_t subsets = (kind == 3) ? 3u : 2u;
	uint64_t err[8];
	for (uint32_t i = 0; i < want; i++) { err[i] = UINT64_MAX; out[i] = 0; }
	BU_ROLL
	for (uint32_t p = 0; p < total; p++)
	{
		const uint64_t e = estimate_pattern_error(T, base + p * 16, subsets, wbits, comps, px);
		for (uint32_t i = 0; i < want; i++)
			if (e < err[i])
			{
				for (uint32_t j = want - 1; j > i; --j) { out[j] = out[j - 1]; err[j] = err[j - 1]; }
				out[i] = (uint8_t)p;
				err[i] = e;
				break;
			}
	}
}

// ---- mode descriptors ---------------------------------------------------------------------------------------------------

BU_FI cell_cfg mode_cell_cfg(const bu_tables* T, uint32_t mode, const level_opts& o, bool cell_has_alpha)
{
	cell_cfg c;
	c.wbits = T->mode_weight_bits[mode];
	c.range = T->mode_endpoint_range[mode];
	c.slot = (uint32_t)T->range_slot[c.range];
	c.has_alpha = cell_has_alpha ? 1u : 0u;
	c.uber = o.uber;
	c.ls_passes = o.ls_passes;
	return c;
}

BU_FI int unq_sum3(const bu_tables* T, uint32_t slot, const uint8_t* ep, int which)
{
	const uint8_t* u = T->unq + slot * 256;
	return (int)u[ep[0 + which]] + (int)u[ep[2 + which]] + (int)u[ep[4 + which]];
}

BU_FI void candidate_clear(candidate& c, uint32_t mode)
{
	c.mode = (uint8_t)mode; c.pattern = 0; c.ccs = 0; c.valid = 1; c.pad = 0; c.uastc_err = 0; c.bc7_err = 0;
	BU_ROLL
	for (int i = 0; i < 18; i++) c.ep[i] = 0;
	BU_ROLL
	for (int i = 0; i < 32; i++) c.w[i] = 0;
}

// Single-subset, single-plane modes 0,1,5,10,12,14,15,18 (uastc_enc.cpp:470, 558, 1140, 1818, 2096, 2335, 2422).
BU_NI inline void gen_one_subset(const bu_tables* T, uint32_t mode, const level_opts& o, const uint32_t* px, candidate& out)
{
	candidate_clear(out, mode);
	const bool rgba = T->mode_comps[mode] == 4, la = T->mode_comps[mode] == 2;
	const cell_cfg cfg = mode_cell_cfg(T, mode, o, rgba || la);
	const uint32_t top = (1u << cfg.wbits) - 1;

	uint32_t src[16];
	BU_ROLL
	for (int i = 0; i < 16; i++) src[i] = la ? px_make(px_c(px[i], 0), 0, 0, px_c(px[i], 3)) : px[i]; // (l,0,0,a): both channels weigh equally

	cell_result r;
	cell_compress(T, cfg, src, 16, r);

	if (la)
	{
		out.ep[0] = r.astc_lo[0]; out.ep[1] = r.astc_hi[0];
		out.ep[2] = r.astc_lo[3]; out.ep[3] = r.astc_hi[3];
		BU_ROLL
		for (int i = 0; i < 16; i++) out.w[i] = r.sel[i];
		return;
	}
	const int nc = rgba ? 4 : 3;
	for (int c = 0; c < nc; c++) { out.ep[c * 2] = r.astc_lo[c]; out.ep[c * 2 + 1] = r.astc_hi[c]; }
	const bool invert = unq_sum3(T, cfg.slot, out.ep, 1) < unq_sum3(T, cfg.slot, out.ep, 0);
	if (invert)
		for (int c = 0; c < nc; c++) { const uint8_t t = out.ep[c * 2]; out.ep[c * 2] = out.ep[c * 2 + 1]; out.ep[c * 2 + 1] = t; }
	BU_ROLL
	for (int i = 0; i < 16; i++) out.w[i] = (uint8_t)(invert ? top - r.sel[i] : r.sel[i]);
}

// Two/three-subset modes 2,3,4,7,9,16 for a given common pattern (uastc_enc.cpp:673, 823, 992, 1357, 1596).
BU_NI inline void gen_multi_subset(const bu_tables* T, uint32_t mode, uint32_t pattern, const level_opts& o, const uint32_t* px, candidate& out)
{
	candidate_clear(out, mode);
	out.pattern = (uint8_t)pattern;
	const uint32_t subsets = T->mode_subsets[mode];
	const bool rgba = (mode == 9), la = (mode == 16);
	const cell_cfg cfg = mode_cell_cfg(T, mode, o, rgba || la);
	const uint32_t top = (1u << cfg.wbits) - 1;
	const uint8_t* part = fit_partition(T, mode, pattern);

	cell_result r[3];
	uint8_t texel_slot[16];
	BU_ROLL
	for (uint32_t s = 0; s < subsets; s++)
	{
		uint32_t sub[16];
		uint32_t n = 0;
		BU_ROLL
		for (int i = 0; i < 16; i++)
			if (part[i] == s)
			{
				texel_slot[i] = (uint8_t)n;
				sub[n++] = la ? px_make(px_c(px[i], 0), 0, 0, px_c(px[i], 3)) : px[i];
			}
		cell_compress(T, cfg, sub, n, r[s]);
	}

	// Map ASTC subset -> fitted (BC7-numbered) subset.
	uint32_t fit_of_astc[3] = { 0, 1, 2 };
	if (mode == 3)
	{
		const uint32_t perm = T->cp3_perm[pattern];
		for (int a = 0; a < 3; a++) fit_of_astc[a] = T->astc_to_bc7_perm[perm * 3 + a];
	}
	else if (mode != 7 && T->cp2_invert[pattern]) { fit_of_astc[0] = 1; fit_of_astc[1] = 0; }

	bool invert[3] = { false, false, false };
	BU_ROLL
	for (uint32_t a = 0; a < subsets; a++)
	{
		const cell_result& q = r[fit_of_astc[a]];
		if (la)
		{
			uint8_t* e = out.ep + a * 4;
			e[0] = q.astc_lo[0]; e[1] = q.astc_hi[0]; e[2] = q.astc_lo[3]; e[3] = q.astc_hi[3];
			continue; // LA endpoints are never reordered here (uastc_enc.cpp:1751-1758)
		}
		const int nc = rgba ? 4 : 3;
		uint8_t* e = out.ep + a * nc * 2;
		for (int c = 0; c < nc; c++) { e[c * 2] = q.astc_lo[c]; e[c * 2 + 1] = q.astc_hi[c]; }
		if (unq_sum3(T, cfg.slot, e, 1) < unq_sum3(T, cfg.slot, e, 0))
		{
			for (int c = 0; c < nc; c++) { const uint8_t t = e[c * 2]; e[c * 2] = e[c * 2 + 1]; e[c * 2 + 1] = t; }
			invert[a] = true;
		}
	}

	BU_ROLL
	for (int i = 0; i < 16; i++)
	{
		const uint32_t f = part[i];
		uint32_t a = 0;
		for (uint32_t k = 0; k < subsets; k++) if (fit_of_astc[k] == f) { a = k; break; }
		const uint32_t s = r[f].sel[texel_slot[i]];
		out.w[i] = (uint8_t)(invert[a] ? top - s : s);
	}
}

// Dual-plane modes 6,11,13,17 for a given rotation / second-plane component (uastc_enc.cpp:1223, 1905, 2182).
BU_NI inline void gen_dual_plane(const bu_tables* T, uint32_t mode, uint32_t rot, const level_opts& o, const uint32_t* px, candidate& out)
{
	candidate_clear(out, mode);
	const cell_cfg cfg = mode_cell_cfg(T, mode, o, false); // both planes are fitted as RGB cells
	const uint32_t top = (1u << cfg.wbits) - 1;

	uint32_t p0[16], p1[16];
	BU_ROLL
	for (int i = 0; i < 16; i++)
	{
		const uint32_t p = px[i];
		if (mode == 17)
		{
			const uint32_t l = px_c(p, 0), a = px_c(p, 3);
			p1[i] = px_make(a, a, a, 255);
			p0[i] = px_make(l, l, l, 255);
		}
		else
		{
			const uint32_t c = px_c(p, rot);
			p1[i] = px_make(c, c, c, 255);
			// mode 6 (RGB source): the rotated channel is neutralised with 255; modes 11/13: alpha moves into the rotated channel.
			p0[i] = (mode == 6) ? px_set(p, rot, 255) : px_set(px_set(p, rot, px_c(p, 3)), 3, 255);
		}
	}

	cell_result r0, r1;
	cell_compress(T, cfg, p0, 16, r0);
	cell_compress(T, cfg, p1, 16, r1);

	bool invert = false;
	if (mode == 17)
	{
		out.ccs = 3;
		out.ep[0] = r0.astc_lo[0]; out.ep[1] = r0.astc_hi[0];
		out.ep[2] = r1.astc_lo[0]; out.ep[3] = r1.astc_hi[0];
	}
	else
	{
		out.ccs = (uint8_t)rot;
		for (uint32_t c = 0; c < 3; c++)
		{
			const cell_result& q = (rot == c) ? r1 : r0;
			out.ep[c * 2] = q.astc_lo[c]; out.ep[c * 2 + 1] = q.astc_hi[c];
		}
		if (mode != 6)
		{
			if (rot == 3) { out.ep[6] = r1.astc_lo[0]; out.ep[7] = r1.astc_hi[0]; }
			else { out.ep[6] = r0.astc_lo[rot]; out.ep[7] = r0.astc_hi[rot]; }
		}
		if (unq_sum3(T, cfg.slot, out.ep, 1) < unq_sum3(T, cfg.slot, out.ep, 0))
		{
			const int nc = (mode == 6) ? 3 : 4;
			for (int c = 0; c < nc; c++) { const uint8_t t = out.ep[c * 2]; out.ep[c * 2] = out.ep[c * 2 + 1]; out.ep[c * 2 + 1] = t; }
			invert = true;
		}
	}
	BU_ROLL
	for (int i = 0; i < 16; i++)
	{
		out.w[i * 2] = (uint8_t)(invert ? top - r0.sel[i] : r0.sel[i]);
		out.w[i * 2 + 1] = (uint8_t)(invert ? top - r1.sel[i] : r1.sel[i]);
	}
}

// ---- decode a candidate to texels (transcoder.cpp:15743 unpack_uastc from an astc_block_desc) ------------------------------

BU_NI inline void decode_candidate(const bu_tables* T, const candidate& c, uint32_t* out)
{
	const uint32_t mode = c.mode;
	const uint32_t subsets = T->mode_subsets[mode], comps = T->mode_comps[mode], planes = T->mode_planes[mode];
	const uint32_t wbits = T->mode_weight_bits[mode];
	const uint8_t* u = T->unq + (uint32_t)T->range_slot[T->mode_endpoint_range[mode]] * 256;
	const uint8_t* wt = T->weights + wbits * 32;

	uint32_t e0[3], e1[3];
	BU_ROLL
	for (uint32_t s = 0; s < subsets; s++)
	{
		const uint8_t* e = c.ep + s * comps * 2;
		if (comps == 2)
		{
			const uint32_t ll = u[e[0]], lh = u[e[1]], al = u[e[2]], ah = u[e[3]];
			e0[s] = px_make(ll, ll, ll, al);
			e1[s] = px_make(lh, lh, lh, ah);
		}
		else
		{
			e0[s] = px_make(u[e[0]], u[e[2]], u[e[4]], comps == 4 ? u[e[6]] : 255u);
			e1[s] = px_make(u[e[1]], u[e[3]], u[e[5]], comps == 4 ? u[e[7]] : 255u);
		}
	}

	const uint8_t* part = (subsets >= 2) ? astc_partition(T, mode, c.pattern) : nullptr;
	BU_ROLL
	for (int i = 0; i < 16; i++)
	{
		const uint32_t s = part ? part[i] : 0;
		const uint32_t w0 = wt[c.w[planes == 2 ? i * 2 : i]];
		const uint32_t w1 = (planes == 2) ? wt[c.w[i * 2 + 1]] : w0;
		uint32_t p = 0;
		for (uint32_t k = 0; k < 4; k++)
		{
			uint32_t v;
			if (comps == 3 && k == 3) v = 255; // RGB modes decode opaque
			else v = astc_lerp(px_c(e0[s], k), px_c(e1[s], k), (planes == 2 && k == c.ccs) ? w1 : w0);
			p |= v << (k * 8);
		}
		out[i] = p;
	}
}

// ---- BC7 transcode of a candidate, decoded to texels -------------------------------------------------------------------
// Composition of transcode_uastc_to_bc7 (transcoder.cpp:16034), encode_bc7_block (14657) and bc7u::unpack_bc7 (30120):
// the encode/unpack pair is the identity on the logical block (anchor flips re-order endpoints and invert selectors
// together), so the logical BC7 block is decoded directly.

struct pbit_fit { uint8_t lo[4], hi[4]; uint32_t p0, p1; };

// determine_unique_pbits (transcoder.cpp:15950): independent p-bit per endpoint.
BU_NI inline void fit_unique_pbits(uint32_t total_comps, uint32_t comp_bits, const float* xl, const float* xh, pbit_fit& f)
{
	const uint32_t total_bits = comp_bits + 1;
	const int iscalep = (1 << total_bits) - 1;
	const float scalep = (float)iscalep;
	float best0 = 1e+9f, best1 = 1e+9f;
	for (int c = 0; c < 4; c++) { f.lo[c] = 0; f.hi[c] = 0; }
	f.p0 = 0; f.p1 = 0;
	for (int p = 0; p < 2; p++)
	{
		uint8_t mn[4], mx[4];
		float err0 = 0, err1 = 0;
		for (uint32_t c = 0; c < 4; c++)
		{
			mn[c] = (uint8_t)clampi(((int)((xl[c] * scalep - (float)p) / 2.0f + .5f)) * 2 + p, p, iscalep - 1 + p);
			mx[c] = (uint8_t)clampi(((int)((xh[c] * scalep - (float)p) / 2.0f + .5f)) * 2 + p, p, iscalep - 1 + p);
		}
		for (uint32_t c = 0; c < total_comps; c++)
		{
			uint8_t sl = (uint8_t)(mn[c] << (8 - total_bits)); sl |= (uint8_t)(sl >> total_bits);
			uint8_t sh = (uint8_t)(mx[c] << (8 - total_bits)); sh |= (uint8_t)(sh >> total_bits);
			const float d0 = (float)sl - xl[c] * 255.0f, d1 = (float)sh - xh[c] * 255.0f;
			err0 += d0 * d0;
			err1 += d1 * d1;
		}
		if (err0 < best0) { best0 = err0; f.p0 = (uint32_t)p; for (int c = 0; c < 4; c++) f.lo[c] = mn[c] >> 1; }
		if (err1 < best1) { best1 = err1; f.p1 = (uint32_t)p; for (int c = 0; c < 4; c++) f.hi[c] = mx[c] >> 1; }
	}
}

// determine_shared_pbits (transcoder.cpp:15897): one p-bit for both endpoints.
BU_NI inline void fit_shared_pbits(uint32_t total_comps, uint32_t comp_bits, const float* xl, const float* xh, pbit_fit& f)
{
	const uint32_t total_bits = comp_bits + 1;
	const int iscalep = (1 << total_bits) - 1;
	const float scalep = (float)iscalep;
	float best = 1e+9f;
	for (int c = 0; c < 4; c++) { f.lo[c] = 0; f.hi[c] = 0; }
	f.p0 = 0; f.p1 = 0;
	for (int p = 0; p < 2; p++)
	{
		uint8_t mn[4], mx[4];
		for (uint32_t c = 0; c < 4; c++)
		{
			mn[c] = (uint8_t)clampi(((int)((xl[c] * scalep - (float)p) / 2.0f + .5f)) * 2 + p, p, iscalep - 1 + p);
			mx[c] = (uint8_t)clampi(((int)((xh[c] * scalep - (float)p) / 2.0f + .5f)) * 2 + p, p, iscalep - 1 + p);
		}
		float err = 0;
		for (uint32_t c = 0; c < total_comps; c++)
		{
			uint8_t sl = (uint8_t)(mn[c] << (8 - total_bits)); sl |= (uint8_t)(sl >> total_bits);
			uint8_t sh = (uint8_t)(mx[c] << (8 - total_bits)); sh |= (uint8_t)(sh >> total_bits);
			const float d0 = ((float)sl / 255.0f) - xl[c], d1 = ((float)sh / 255.0f) - xh[c];
			err += d0 * d0 + d1 * d1;
		}
		if (err < best)
		{
			best = err; f.p0 = f.p1 = (uint32_t)p;
			for (int c = 0; c < 4; c++) { f.lo[c] = mn[c] >> 1; f.hi[c] = mx[c] >> 1; }
		}
	}
}

BU_FI uint32_t bc7_dq_p(uint32_t v, uint32_t pbit, uint32_t bits) { const uint32_t tb = bits + 1; v = ((v << 1) | pbit) << (8 - tb); return v | (v >> tb); }
BU_FI uint32_t bc7_dq(uint32_t v, uint32_t bits) { v <<= (8 - bits); return v | (v >> bits); }
BU_FI uint32_t bc7_lerp(uint32_t l, uint32_t h, uint32_t w) { return (l * (64 - w) + h * w + 32) >> 6; }

// BC7 mode 6 / mode 1 weights are 4-bit; UASTC modes with 5- or 3-bit weights map onto them (transcoder.cpp:14600-14650).
BU_TABLE(uint8_t, bc7_weight5_to_4, [32], { 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 6, 7, 8, 9, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 14, 14, 15, 15 })
BU_TABLE(uint8_t, bc7_weight3_to_4, [8], { 0, 2, 4, 6, 9, 11, 13, 15 })

// The BC7 endpoints of the transcode depend on the UASTC endpoints (and pattern / rotation) only, the texels on those endpoints
// and the weights: the two halves are separate so that uastc_rdo's trials, which splice weights only, fit the p-bits once.
struct bc7_endpoints { uint32_t l[3][4], h[3][4]; }; // [BC7 subset][channel]

BU_NI inline void bc7_endpoints_of(const bu_tables* T, const candidate& c, bc7_endpoints& E)
{
	const uint32_t mode = c.mode;
	const uint32_t range = T->mode_endpoint_range[mode];
	const uint8_t* u = T->unq + (uint32_t)T->range_slot[range] * 256;
	const uint32_t comps = T->mode_comps[mode];

	switch (mode)
	{
	case 0: case 5: case 10: case 12: case 14: case 15: case 18:
	{
		// -> BC7 mode 6: 7777.1 endpoints, 4-bit indices
		float xl[4], xh[4];
		if (comps == 2)
		{
			xl[0] = (float)u[c.ep[0]] / 255.0f; xh[0] = (float)u[c.ep[1]] / 255.0f;
			xl[1] = xl[2] = xl[0]; xh[1] = xh[2] = xh[0];
			xl[3] = (float)u[c.ep[2]] / 255.0f; xh[3] = (float)u[c.ep[3]] / 255.0f;
		}
		else
		{
			for (int k = 0; k < 3; k++) { xl[k] = (float)u[c.ep[k * 2]] / 255.0f; xh[k] = (float)u[c.ep[k * 2 + 1]] / 255.0f; }
			if (comps == 4) { xl[3] = (float)u[c.ep[6]] / 255.0f; xh[3] = (float)u[c.ep[7]] / 255.0f; }
			else { xl[3] = 1.0f; xh[3] = 1.0f; }
		}
		pbit_fit f;
		fit_unique_pbits(comps == 2 ? 4 : comps, 7, xl, xh, f);
		if (comps == 3) { f.lo[3] = 127; f.hi[3] = 127; }
		for (int k = 0; k < 4; k++) { E.l[0][k] = ((uint32_t)f.lo[k] << 1) | f.p0; E.h[0][k] = ((uint32_t)f.hi[k] << 1) | f.p1; }
		return;
	}
	case 1: case 4:
	{
		// -> BC7 mode 3: two subsets, 777.1 endpoints with unique p-bits, 2-bit indices. Mode 1 replicates one endpoint pair.
		const uint32_t subsets = (mode == 1) ? 1u : 2u;
		const bool inv = (mode == 4) && T->cp2_invert[c.pattern];
		BU_ROLL
		for (uint32_t s = 0; s < subsets; s++)
		{
			float xl[4], xh[4];
			for (int k = 0; k < 3; k++)
			{
				// mode 1 (range 20) uses the raw 8-bit values (transcoder.cpp:16150); range 20 unquantises to itself anyway
				xl[k] = (float)u[c.ep[k * 2 + s * 6]] / 255.0f;
				xh[k] = (float)u[c.ep[k * 2 + 1 + s * 6]] / 255.0f;
			}
			xl[3] = 1.0f; xh[3] = 1.0f;
			pbit_fit f;
			fit_unique_pbits(3, 7, xl, xh, f);
			const uint32_t b = inv ? 1 - s : s;
			for (int k = 0; k < 3; k++) { E.l[b][k] = ((uint32_t)f.lo[k] << 1) | f.p0; E.h[b][k] = ((uint32_t)f.hi[k] << 1) | f.p1; }
		}
		return;
	}
	case 2:
	{
		// -> BC7 mode 1: two subsets, 666 endpoints with a shared p-bit, 3-bit indices
		const bool inv = T->cp2_invert[c.pattern] != 0;
		BU_ROLL
		for (uint32_t s = 0; s < 2; s++)
		{
			float xl[4], xh[4];
			for (int k = 0; k < 3; k++)
			{
				uint32_t v = c.ep[k * 2 + s * 6]; v = (v << 4) | v; xl[k] = (float)v / 255.0f;
				v = c.ep[k * 2 + s * 6 + 1]; v = (v << 4) | v; xh[k] = (float)v / 255.0f;
			}
			xl[3] = 1.0f; xh[3] = 1.0f;
			pbit_fit f;
			fit_shared_pbits(3, 6, xl, xh, f);
			const uint32_t b = inv ? 1 - s : s;
			for (int k = 0; k < 3; k++) { E.l[b][k] = bc7_dq_p(f.lo[k], f.p0, 6); E.h[b][k] = bc7_dq_p(f.hi[k], f.p0, 6); }
		}
		return;
	}
	case 3: case 7:
	{
		// -> BC7 mode 2: three subsets, 555 endpoints, 2-bit indices
		if (mode == 3)
		{
			const uint32_t perm = T->cp3_perm[c.pattern];
			BU_ROLL
			for (uint32_t s = 0; s < 3; s++)
			{
				const uint32_t b = T->astc_to_bc7_perm[perm * 3 + s];
				for (int k = 0; k < 3; k++)
				{
					E.l[b][k] = bc7_dq(((uint32_t)u[c.ep[k * 2 + s * 6]] * 31 + 127) / 255, 5);
					E.h[b][k] = bc7_dq(((uint32_t)u[c.ep[k * 2 + 1 + s * 6]] * 31 + 127) / 255, 5);
				}
			}
		}
		else
		{
			const uint32_t kk = T->cp73_k[c.pattern];
			BU_ROLL
			for (uint32_t b = 0; b < 3; b++)
			{
				const uint32_t s = bc7_part3_to_2(b, kk);
				for (int k = 0; k < 3; k++)
				{
					E.l[b][k] = bc7_dq(((uint32_t)u[c.ep[k * 2 + s * 6]] * 31 + 127) / 255, 5);
					E.h[b][k] = bc7_dq(((uint32_t)u[c.ep[k * 2 + 1 + s * 6]] * 31 + 127) / 255, 5);
				}
			}
		}
		return;
	}
	case 6: case 11: case 13: case 17:
	{
		// -> BC7 mode 5: 777 colour + 8-bit alpha, separate 2-bit index planes, channel rotation
		uint32_t l[4], h[4];
		if (comps == 2)
		{
			l[0] = ((uint32_t)u[c.ep[0]] * 127 + 127) / 255; h[0] = ((uint32_t)u[c.ep[1]] * 127 + 127) / 255;
			l[1] = l[2] = l[0]; h[1] = h[2] = h[0];
			l[3] = u[c.ep[2]]; h[3] = u[c.ep[3]];
		}
		else
		{
			BU_ROLL
			for (uint32_t a = 0; a < 4; a++)
			{
				uint32_t b = a;
				if (a == c.ccs) b = 3; else if (a == 3) b = c.ccs;
				uint32_t lv = 255, hv = 255;
				if (a < comps) { lv = u[c.ep[a * 2]]; hv = u[c.ep[a * 2 + 1]]; }
				if (b < 3) { lv = (lv * 127 + 127) / 255; hv = (hv * 127 + 127) / 255; }
				l[b] = lv; h[b] = hv;
			}
		}
		for (int k = 0; k < 3; k++) { l[k] = bc7_dq(l[k], 7); h[k] = bc7_dq(h[k], 7); }
		for (int k = 0; k < 4; k++) { E.l[0][k] = l[k]; E.h[0][k] = h[k]; }
		return;
	}
	default: // 9, 16
	{
		// -> BC7 mode 7: two subsets, 5555.1 endpoints with unique p-bits, 2-bit indices
		const bool inv = T->cp2_invert[c.pattern] != 0;
		BU_ROLL
		for (uint32_t s = 0; s < 2; s++)
		{
			float xl[4], xh[4];
			if (comps == 2)
			{
				xl[0] = (float)u[c.ep[0 + s * 4]] / 255.0f; xh[0] = (float)u[c.ep[1 + s * 4]] / 255.0f;
				xl[1] = xl[2] = xl[0]; xh[1] = xh[2] = xh[0];
				xl[3] = (float)u[c.ep[2 + s * 4]] / 255.0f; xh[3] = (float)u[c.ep[3 + s * 4]] / 255.0f;
			}
			else
				for (int k = 0; k < 4; k++) { xl[k] = (float)u[c.ep[k * 2 + s * 8]] / 255.0f; xh[k] = (float)u[c.ep[k * 2 + 1 + s * 8]] / 255.0f; }
			pbit_fit f;
			fit_unique_pbits(4, 5, xl, xh, f);
			const uint32_t b = inv ? 1 - s : s;
			for (int k = 0; k < 4; k++) { E.l[b][k] = bc7_dq_p(f.lo[k], f.p0, 5); E.h[b][k] = bc7_dq_p(f.hi[k], f.p1, 5); }
		}
		return;
	}
	}
}

BU_NI inline void bc7_texels(const bu_tables* T, const candidate& c, const bc7_endpoints& E, uint32_t* out)
{
	const uint32_t mode = c.mode;
	switch (mode)
	{
	case 0: case 5: case 10: case 12: case 14: case 15: case 18:
	{
		const uint32_t* l = E.l[0];
		const uint32_t* h = E.h[0];
		BU_ROLL
		for (int i = 0; i < 16; i++)
		{
			uint32_t s = c.w[i];
			if (mode == 18) s = BU_TABLE_REF(bc7_weight5_to_4)[s];
			else if (mode == 14) s = s * 5;
			else if (mode == 5 || mode == 12) s = BU_TABLE_REF(bc7_weight3_to_4)[s];
			const uint32_t w = T->bc7_weights4[s];
			out[i] = px_make(bc7_lerp(l[0], h[0], w), bc7_lerp(l[1], h[1], w), bc7_lerp(l[2], h[2], w), bc7_lerp(l[3], h[3], w));
		}
		return;
	}
	case 1: case 4:
	{
		const uint8_t* part = (mode == 4) ? T->bc7_part2 + c.pattern * 16 : nullptr;
		BU_ROLL
		for (int i = 0; i < 16; i++)
		{
			const uint32_t b = part ? part[i] : 0;
			const uint32_t w = T->weights[2 * 32 + c.w[i]];
			out[i] = px_make(bc7_lerp(E.l[b][0], E.h[b][0], w), bc7_lerp(E.l[b][1], E.h[b][1], w), bc7_lerp(E.l[b][2], E.h[b][2], w), 255);
		}
		return;
	}
	case 2:
	{
		const uint8_t* part = T->bc7_part2 + c.pattern * 16;
		BU_ROLL
		for (int i = 0; i < 16; i++)
		{
			const uint32_t b = part[i];
			const uint32_t w = T->weights[3 * 32 + c.w[i]];
			out[i] = px_make(bc7_lerp(E.l[b][0], E.h[b][0], w), bc7_lerp(E.l[b][1], E.h[b][1], w), bc7_lerp(E.l[b][2], E.h[b][2], w), 255);
		}
		return;
	}
	case 3: case 7:
	{
		const uint8_t* part = (mode == 3) ? T->bc7_part3 + c.pattern * 16 : T->bc7_part73 + c.pattern * 16;
		BU_ROLL
		for (int i = 0; i < 16; i++)
		{
			const uint32_t b = part[i];
			const uint32_t w = T->weights[2 * 32 + c.w[i]];
			out[i] = px_make(bc7_lerp(E.l[b][0], E.h[b][0], w), bc7_lerp(E.l[b][1], E.h[b][1], w), bc7_lerp(E.l[b][2], E.h[b][2], w), 255);
		}
		return;
	}
	case 6: case 11: case 13: case 17:
	{
		const uint32_t* l = E.l[0];
		const uint32_t* h = E.h[0];
		const uint32_t rotation = ((uint32_t)c.ccs + 1) & 3;
		BU_ROLL
		for (int i = 0; i < 16; i++)
		{
			uint32_t s0 = c.w[i * 2], s1 = c.w[i * 2 + 1];
			if (mode == 13) { s0 = s0 ? 3 : 0; s1 = s1 ? 3 : 0; }
			const uint32_t w0 = T->weights[2 * 32 + s0], w1 = T->weights[2 * 32 + s1];
			uint32_t v[4] = { bc7_lerp(l[0], h[0], w0), bc7_lerp(l[1], h[1], w0), bc7_lerp(l[2], h[2], w0), bc7_lerp(l[3], h[3], w1) };
			if (rotation >= 1) { const uint32_t t = v[3]; v[3] = v[rotation - 1]; v[rotation - 1] = t; }
			out[i] = px_make(v[0], v[1], v[2], v[3]);
		}
		return;
	}
	default: // 9, 16
	{
		const uint8_t* part = T->bc7_part2 + c.pattern * 16;
		BU_ROLL
		for (int i = 0; i < 16; i++)
		{
			const uint32_t b = part[i];
			const uint32_t w = T->weights[2 * 32 + c.w[i]];
			out[i] = px_make(bc7_lerp(E.l[b][0], E.h[b][0], w), bc7_lerp(E.l[b][1], E.h[b][1], w), bc7_lerp(E.l[b][2], E.h[b][2], w), bc7_lerp(E.l[b][3], E.h[b][3], w));
		}
		return;
	}
	}
}

BU_HD inline void decode_bc7_transcode(const bu_tables* T, const candidate& c, uint32_t* out)
{
	bc7_endpoints E;
	bc7_endpoints_of(T, c, E);
	bc7_texels(T, c, E, out);
}

// ---- candidate scoring (uastc_enc.cpp:2510 compute_block_error + 3471-3488 metric selection) --------------------------------

BU_HD inline uint32_t block_error(const uint32_t* src, const uint32_t* dec, block_class k)
{
	uint32_t er = 0, eg = 0, eb = 0, ea = 0;
	BU_ROLL
	for (int i = 0; i < 16; i++)
	{
		er += sq_diff((int)px_c(src[i], 0), (int)px_c(dec[i], 0));
		eg += sq_diff((int)px_c(src[i], 1), (int)px_c(dec[i], 1));
		eb += sq_diff((int)px_c(src[i], 2), (int)px_c(dec[i], 2));
		ea += sq_diff((int)px_c(src[i], 3), (int)px_c(dec[i], 3));
	}
	if (k.is_la) return er + ea;
	if (k.has_alpha) return er + eg + eb + ea;
	return er + eg + eb;
}

BU_HD inline void score_candidate(const bu_tables* T, const uint32_t* src, block_class k, candidate& c)
{
	uint32_t dec[16];
	decode_candidate(T, c, dec);
	c.uastc_err = block_error(src, dec, k);
	decode_bc7_transcode(T, c, dec);
	c.bc7_err = block_error(src, dec, k);
}

// ---- final choice among a block's candidates (uastc_enc.cpp:3397-3549) ------------------------------------------------------
// errs: (uastc_err, bc7_err) pairs in candidate order; modes: UASTC mode per candidate. Returns the chosen index.

BU_NI inline int select_candidate(uint32_t n, const uint32_t* uastc_err, const uint32_t* bc7_err, const uint8_t* modes, uint32_t flags)
{
	if (n == 1) return 0;
	const bool favor_uastc = (flags & 8) != 0;              // cPackUASTCFavorUASTCError
	const bool favor_bc7 = !favor_uastc && ((flags & 16) != 0); // cPackUASTCFavorBC7Error
	const bool favor_simple = (flags & 512) != 0;           // cPackUASTCFavorSimplerModes
	const uint64_t bc7_w = favor_bc7 ? 100 : (favor_uastc ? 0 : 50);
	const uint64_t uastc_w = favor_bc7 ? 0 : 100;

	double best_f = 1e+20f;
	BU_ROLL
	for (uint32_t i = 0; i < n; i++)
	{
		const uint64_t overall = ((uint64_t)bc7_err[i] * bc7_w) / 100 + ((uint64_t)uastc_err[i] * uastc_w) / 100;
		if (!overall) return (int)i;
		const float f = sqrtf((float)uastc_err[i]);
		if ((double)f < best_f) best_f = f;
	}

	int best = -1;
	uint64_t best_err = UINT64_MAX;
	const bool all = (best_f == 0.0) || favor_bc7;
	BU_ROLL
	for (uint32_t i = 0; i < n; i++)
	{
		if (!all)
		{
			const double delta = (double)sqrtf((float)uastc_err[i]) / best_f;
			if (!(delta <= (double)1.3f)) continue;
		}
		const uint64_t overall = ((uint64_t)bc7_err[i] * bc7_w) / 100 + ((uint64_t)uastc_err[i] * uastc_w) / 100;
		const float weight = (favor_simple && (modes[i] == 0 || modes[i] == 10)) ? .8f : 1.0f;
		const uint64_t w = (uint64_t)((float)overall * weight);
		if (w < best_err)
		{
			best_err = w;
			best = (int)i;
			if (!best_err) break;
		}
	}
	return best;
}

} // namespace bu
