// bu_rdo.h -- UASTC RDO post-pass building blocks (reference: uastc_rdo / uastc_rdo_blocks, encoder/basisu_uastc_enc.cpp:3824-4164).
//
// The reference walks a chunk of blocks in order; for each block it tries to replace the block's selector (weight) bits by
// the bits of each of the previous <= 256 blocks, keeps the replacement minimising  ms_err * smooth_scale + lz_bits * lambda,
// optionally refits mode-0 endpoints to the new selectors, and recomputes the transcode hints.  The chain is inherently
// sequential (block i reads the rewritten bits of blocks i-256..i-1), but
//   * the <= 256 trials of one step are independent  -> one thread each, merged by an (cost, distance) arg-min;
//   * no later decision reads a block's hint bits (every mode's selector field starts at bit >= 49, hints end before bit 31),
//     so the hint recomputation -- the most expensive part -- is deferred to a fully parallel pass over the modified blocks.
// Everything here is bit-exact with the reference, including its float cost arithmetic.
#pragma once
#include "bu_slots.h"

namespace bu {

struct rdo_params
{
	uint32_t lz_dict_size;
	float lambda, max_allowed_rms_increase_ratio, skip_block_rms_thresh;
	uint32_t endpoint_refinement;
	float max_smooth_block_std_dev, smooth_block_max_error_scale;
	uint32_t lz_literal_cost;
};

struct block_bits { uint64_t lo, hi; };

BU_FI uint64_t bits_read(const block_bits& b, uint32_t ofs, uint32_t n) // n <= 64, ofs + n <= 128
{
	if (!n) return 0;
	uint64_t v;
	if (ofs >= 64) v = b.hi >> (ofs - 64);
	else { v = b.lo >> ofs; if (ofs && ofs + n > 64) v |= b.hi << (64 - ofs); }
	return (n == 64) ? v : (v & ((1ull << n) - 1));
}
BU_FI void bits_write(block_bits& b, uint32_t ofs, uint32_t n, uint64_t v) // overwrite n bits at ofs
{
	if (!n) return;
	const uint64_t mask = (n == 64) ? ~0ull : ((1ull << n) - 1);
	v &= mask;
	if (ofs >= 64) { b.hi = (b.hi & ~(mask << (ofs - 64))) | (v << (ofs - 64)); return; }
	b.lo = (b.lo & ~(mask << ofs)) | (v << ofs);
	if (ofs && ofs + n > 64) { const uint32_t sh = 64 - ofs; b.hi = (b.hi & ~(mask >> sh)) | (v >> sh); }
}

// g_uastc_mode_selector_bits (uastc_enc.cpp:3729): first bit and length of each mode's weight field.
BU_TABLE(uint8_t, selector_field_first, [19], { 65, 69, 73, 89, 89, 68, 66, 89, 0, 97, 65, 66, 81, 94, 92, 62, 98, 61, 49 })
BU_TABLE(uint8_t, selector_field_total, [19], { 63, 31, 46, 29, 30, 47, 62, 30, 0, 30, 63, 62, 47, 30, 31, 63, 30, 62, 79 })
BU_FI void mode_selector_field(uint32_t mode, uint32_t& first, uint32_t& total)
{
	first = BU_TABLE_REF(selector_field_first)[mode]; total = BU_TABLE_REF(selector_field_total)[mode];
}

// compute_match_cost_estimate (uastc_enc.cpp:3775). The two miniz distance-extra-bits tables it indexes are
// floor(log2(dist)) - 1 for 4 <= dist < 512 (0 below 4) and floor(log2(dist >> 8)) + 7 for 512 <= dist < 32768.
BU_FI uint32_t ilog2u(uint32_t v) { uint32_t r = 0; while (v >>= 1) r++; return r; }
BU_FI uint32_t match_cost_estimate(uint32_t dist)
{
	uint32_t dist_cost = 5;
	if (dist < 512) dist_cost += (dist < 4) ? 0 : ilog2u(dist) - 1;
	else
	{
		const uint32_t d = (dist < 32767 ? dist : 32767) >> 8;
		dist_cost += ilog2u(d) + 7;
		while (dist >= 32768) { dist_cost++; dist >>= 1; }
	}
	return 7 + dist_cost;
}

// unpack_uastc(blk, unpacked, blue_contract_check = false, ...) (transcoder.cpp:15282): mode, pattern, CCS, endpoints, weights.
// Returns false for an invalid mode/pattern. Solid-colour blocks return true with c.mode == 8 and nothing else filled.
BU_NI inline bool unpack_block_bits(const bu_tables* T, const block_bits& b, candidate& c)
{
	uint32_t mode = 0xFF;
	const uint32_t first7 = (uint32_t)(b.lo & 127);
	for (uint32_t m = 0; m < 19; m++)
	{
		const uint32_t code = T->mode_huff[m * 2], len = T->mode_huff[m * 2 + 1];
		if ((first7 & ((1u << len) - 1)) == code) { mode = m; break; }
	}
	if (mode == 0xFF) return false;
	candidate_clear(c, mode);
	if (mode == 8) return true;

	uint32_t ofs = T->mode_huff[mode * 2 + 1];
	ofs += T->mode_has_bc1_hint0[mode] + T->mode_has_bc1_hint1[mode] + 8 + (T->mode_has_etc1_bias[mode] ? 5 : 0) + (T->mode_has_alpha[mode] ? 8 : 0);

	const uint32_t subsets = T->mode_subsets[mode], planes = T->mode_planes[mode], comps = T->mode_comps[mode];
	if (subsets == 2) { c.pattern = (uint8_t)bits_read(b, ofs, 5); ofs += 5; if (c.pattern >= ((mode == 7) ? 19 : 30)) return false; }
	else if (subsets == 3) { c.pattern = (uint8_t)bits_read(b, ofs, 4); ofs += 4; if (c.pattern >= 11) return false; }
	if (mode == 6 || mode == 11 || mode == 13) { c.ccs = (uint8_t)bits_read(b, ofs, 2); ofs += 2; }
	else if (mode == 17) c.ccs = 3;

	// endpoints: BISE groups then plain bits (transcoder.cpp:15456-15533)
	const uint32_t total_values = comps * 2 * subsets;
	const uint32_t range = T->mode_endpoint_range[mode];
	const uint32_t ep_bits = T->bise[range * 3], ep_trits = T->bise[range * 3 + 1], ep_quints = T->bise[range * 3 + 2];
	uint32_t tq[8];
	uint32_t total_tqs = 0, bundle = 0, mul = 0;
	if (ep_trits) { total_tqs = (total_values + 4) / 5; bundle = 5; mul = 3; }
	else if (ep_quints) { total_tqs = (total_values + 2) / 3; bundle = 3; mul = 5; }
	for (uint32_t i = 0; i < total_tqs; i++)
	{
		uint32_t nb = ep_trits ? 8 : 7;
		if (i == total_tqs - 1)
		{
			const uint32_t rem = total_values - (total_tqs - 1) * bundle;
			if (ep_trits) { if (rem == 1) nb = 2; else if (rem == 2) nb = 4; else if (rem == 3) nb = 5; else if (rem == 4) nb = 7; }
			else { if (rem == 1) nb = 3; else if (rem == 2) nb = 5; }
		}
		tq[i] = (uint32_t)bits_read(b, ofs, nb); ofs += nb;
	}
	uint32_t accum = 0, left = 0, next = 0;
	for (uint32_t i = 0; i < total_values; i++)
	{
		uint32_t v = (uint32_t)bits_read(b, ofs, ep_bits); ofs += ep_bits;
		if (total_tqs)
		{
			if (!left) { accum = tq[next++]; left = bundle; }
			v |= (accum % mul) << ep_bits;
			accum /= mul; left--;
		}
		c.ep[i] = (uint8_t)v;
	}

	// weights: anchors are stored one bit short (transcoder.cpp:15577-15690)
	const uint32_t wbits = T->mode_weight_bits[mode];
	const uint8_t zero3[3] = { 0, 0, 0 };
	const uint8_t* anchors = (subsets >= 2) ? astc_anchors(T, mode, c.pattern) : zero3;
	const uint32_t plane_shift = planes - 1;
	for (uint32_t i = 0; i < 16 * planes; i++)
	{
		uint32_t nb = wbits;
		for (uint32_t s = 0; s < subsets; s++) if (anchors[s] == (i >> plane_shift)) { nb--; break; }
		c.w[i] = (uint8_t)bits_read(b, ofs, nb); ofs += nb;
	}
	return true;
}

// basist::unpack_uastc(blk, pPixels, srgb = false) (transcoder.cpp:15886): a packed block to its 16 texels.
BU_HD inline bool unpack_block_texels(const bu_tables* T, const block_bits& b, uint32_t* px)
{
	candidate c;
	if (!unpack_block_bits(T, b, c)) { for (int k = 0; k < 16; k++) px[k] = 0; return false; }
	if (c.mode == 8)
	{
		// solid colour: the R, G, B, A bytes follow the mode code (pack_uastc, uastc_enc.cpp:110; transcoder.cpp:15316)
		const uint32_t colour = (uint32_t)bits_read(b, T->mode_huff[8 * 2 + 1], 32);
		for (int k = 0; k < 16; k++) px[k] = colour;
		return true;
	}
	decode_candidate(T, c, px);
	return true;
}

// Weights of `c`'s mode parsed from a 128-bit block whose weight field has been spliced in (the trial of uastc_enc.cpp:3959-3971).
BU_FI void read_weight_field(const bu_tables* T, const candidate& c, const block_bits& b, uint8_t* w)
{
	uint32_t first, total;
	mode_selector_field(c.mode, first, total);
	const uint32_t subsets = T->mode_subsets[c.mode], planes = T->mode_planes[c.mode], wbits = T->mode_weight_bits[c.mode];
	const uint8_t zero3[3] = { 0, 0, 0 };
	const uint8_t* anchors = (subsets >= 2) ? astc_anchors(T, c.mode, c.pattern) : zero3;
	const uint32_t plane_shift = planes - 1;
	uint32_t ofs = first;
	for (uint32_t i = 0; i < 16 * planes; i++)
	{
		uint32_t nb = wbits;
		for (uint32_t s = 0; s < subsets; s++) if (anchors[s] == (i >> plane_shift)) { nb--; break; }
		w[i] = (uint8_t)bits_read(b, ofs, nb); ofs += nb;
	}
}

// (uastc_err + bc7_err) / 2 with the RGBA metric (uastc_enc.cpp:3869-3888).
// E: the BC7 endpoints of the block's transcode (bc7_endpoints_of). They depend on the UASTC endpoints only, so the trials of
// one step, which change weights only, share the current block's.
BU_NI inline uint64_t rdo_block_error(const bu_tables* T, const candidate& c, const bc7_endpoints& E, const uint32_t* px)
{
	uint32_t dec[16];
	uint64_t ue = 0, be = 0;
	decode_candidate(T, c, dec);
	for (int i = 0; i < 16; i++) ue += dist_rgba(px[i], dec[i]);
	bc7_texels(T, c, E, dec);
	for (int i = 0; i < 16; i++) be += dist_rgba(px[i], dec[i]);
	return (ue + be) / 2;
}

// smooth_block_error_scale (uastc_enc.cpp:3847-3861)
BU_HD inline float rdo_smooth_scale(const rdo_params& p, const uint32_t* px)
{
	float max_sd = 0.0f;
	for (uint32_t c = 0; c < 4; c++)
	{
		int64_t total = 0, total2 = 0;
		for (int i = 0; i < 16; i++) { const int v = (int)px_c(px[i], c); total += v; total2 += v * v; }
		const float sd = sqrtf((float)(16 * total2 - total * total)) / (float)16u;
		// maximum(maximum(maximum(r, g), b), a) with a > b ? a : b semantics
		max_sd = (c == 0) ? sd : ((max_sd > sd) ? max_sd : sd);
	}
	float yl = max_sd / p.max_smooth_block_std_dev;
	yl = (yl < 0.0f) ? 0.0f : ((yl > 1.0f) ? 1.0f : yl);
	yl = yl * yl;
	return p.smooth_block_max_error_scale + (1.0f - p.smooth_block_max_error_scale) * yl; // lerp(a, b, t) = a + (b - a) * t
}

// Per-block quantities every trial of the step shares.
struct rdo_step
{
	candidate cur;            // unpacked current block
	bc7_endpoints bc7;        // its BC7-transcode endpoints (shared by every trial of the step)
	block_bits bits;          // its 128 bits
	uint32_t first_sel_bit, total_sel_bits;
	uint64_t cur_sel_bits;
	float cur_ms_err, cur_rms_err, smooth_scale;
};

// One trial: splice `prev`'s weight field into the current block and cost it. match_index = index of the latest block that
// registered these selector bits (or prev_index if none). Returns false if the trial is skipped or rejected.
BU_NI inline bool rdo_trial(const bu_tables* T, const rdo_params& p, const rdo_step& st, const uint32_t* px, const block_bits& prev, int prev_index, int match_index,
	int block_index, float& t_out, block_bits& trial_bits)
{
	if (match_index > prev_index) return false; // this bit pattern is examined at its most recent occurrence (uastc_enc.cpp:3951)
	trial_bits = st.bits;
	const uint32_t n0 = st.total_sel_bits < 64 ? st.total_sel_bits : 64;
	bits_write(trial_bits, st.first_sel_bit, n0, bits_read(prev, st.first_sel_bit, n0));
	if (st.total_sel_bits > 64) bits_write(trial_bits, st.first_sel_bit + 64, st.total_sel_bits - 64, bits_read(prev, st.first_sel_bit + 64, st.total_sel_bits - 64));
	candidate tc = st.cur;
	read_weight_field(T, tc, trial_bits, tc.w);
	const uint64_t err = rdo_block_error(T, tc, st.bc7, px);
	const float ms = (float)err * (1.0f / 64.0f);
	const float rms = sqrtf(ms);
	if (rms > st.cur_rms_err * p.max_allowed_rms_increase_ratio) return false;
	const int match_bits = (int)match_cost_estimate((uint32_t)((block_index - match_index) * 16));
	t_out = ms * st.smooth_scale + (float)match_bits * p.lambda;
	return true;
}

// Mode-0 endpoint refit for fixed selectors (uastc_enc.cpp:4027-4079): astc_mode0_or_18(0, ..., pForce_selectors) with
// 1 least-squares pass and uber level 0. Returns true and updates c.ep if the UASTC error strictly decreases.
BU_NI inline bool rdo_refine_mode0(const bu_tables* T, candidate& c, const uint32_t* px)
{
	uint32_t dec[16];
	decode_candidate(T, c, dec);
	uint64_t best = 0;
	for (int i = 0; i < 16; i++) best += dist_rgba(px[i], dec[i]);

	level_opts o = make_level_opts(0);
	o.uber = 0; o.ls_passes = 1;
	const cell_cfg cfg = mode_cell_cfg(T, 0, o, false);
	cell_result r;
	cell_compress(T, cfg, px, 16, r, c.w);
	candidate tc = c;
	for (int k = 0; k < 3; k++) { tc.ep[k * 2] = r.astc_lo[k]; tc.ep[k * 2 + 1] = r.astc_hi[k]; }
	decode_candidate(T, tc, dec);
	uint64_t trial = 0;
	for (int i = 0; i < 16; i++) trial += dist_rgba(px[i], dec[i]);
	if (trial < best) { c = tc; return true; }
	return false;
}

// Packs a candidate with neutral hint fields (the deferred hint pass rewrites them; field widths depend on the mode only).
BU_NI inline block_bits pack_without_hints(const bu_tables* T, const candidate& c)
{
	uint8_t ep[18], w[32], out[16];
	canonicalize(T, c, ep, w);
	etc1_hint h; h.err = 0; h.order = 0; h.flip = 0; h.diff = 0; h.inten0 = 0; h.inten1 = 0; h.bias = 0;
	pack_block(T, c, ep, w, h, 0, 1, false, false, out);
	block_bits b; b.lo = 0; b.hi = 0;
	for (int i = 0; i < 8; i++) { b.lo |= (uint64_t)out[i] << (i * 8); b.hi |= (uint64_t)out[8 + i] << (i * 8); }
	return b;
}

} // namespace bu
