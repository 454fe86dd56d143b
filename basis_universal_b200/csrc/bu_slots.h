// bu_slots.h -- the candidate-slot schedule of a UASTC effort level, and the per-block sequence that turns slots into a
// packed block.  A slot is one (mode, variant) pair the reference's encode_uastc would append to its results[] array
// (encoder/basisu_uastc_enc.cpp:3289-3350), listed in exactly that order because the index is the final tie-breaker
// (3516, 3538).  variant = plane rotation for dual-plane modes, partition *rank* (into the block's ranked pattern list) when
// partitions are estimated, or the pattern index itself when they are searched exhaustively (level 4).
#pragma once
#include "bu_uastc_pack.h"

namespace bu {

enum { KLASS_LA = 0, KLASS_RGB = 1, KLASS_ALPHA = 2 };
enum { MAX_SLOTS = 176 };

struct slot_desc { uint8_t mode, variant, klass, pad; };

// Indices into block_ranks::r
enum { RANK_M2 = 0, RANK_M3 = 1, RANK_M4 = 2, RANK_M7 = 3, RANK_M9 = 4, RANK_M16 = 8 };
struct block_ranks { uint8_t r[12]; };

BU_HD inline uint32_t build_slots(const level_opts& o, slot_desc* s)
{
	uint32_t n = 0;
	const uint32_t m = o.mode_mask;
	const bool est = o.estimate_partition != 0;
#define BU_ADD(mode_, var_, kl_) do { s[n].mode = (uint8_t)(mode_); s[n].variant = (uint8_t)(var_); s[n].klass = (uint8_t)(kl_); s[n].pad = 0; n++; } while (0)
	// LA modes
	if (m & (1u << 15)) BU_ADD(15, 0, KLASS_LA);
	if (m & (1u << 16)) for (uint32_t v = 0; v < (est ? 4u : 30u); v++) BU_ADD(16, v, KLASS_LA);
	if (m & (1u << 17)) BU_ADD(17, 0, KLASS_LA);
	// RGB modes
	if (m & (1u << 0)) BU_ADD(0, 0, KLASS_RGB);
	if (m & (1u << 1)) BU_ADD(1, 0, KLASS_RGB);
	if (m & (1u << 2)) for (uint32_t v = 0; v < (est ? 1u : 30u); v++) BU_ADD(2, v, KLASS_RGB);
	if (m & (1u << 3)) for (uint32_t v = 0; v < (est ? 1u : 11u); v++) BU_ADD(3, v, KLASS_RGB);
	if (m & (1u << 4)) for (uint32_t v = 0; v < (est ? 1u : 30u); v++) BU_ADD(4, v, KLASS_RGB);
	if (m & (1u << 5)) BU_ADD(5, 0, KLASS_RGB);
	if (m & (1u << 6)) for (uint32_t v = 0; v < 3; v++) BU_ADD(6, v, KLASS_RGB);
	if (m & (1u << 7)) for (uint32_t v = 0; v < (est ? 1u : 19u); v++) BU_ADD(7, v, KLASS_RGB);
	if (m & (1u << 18)) BU_ADD(18, 0, KLASS_RGB);
	// alpha modes
	if (m & (1u << 9)) for (uint32_t v = 0; v < (est ? 4u : 30u); v++) BU_ADD(9, v, KLASS_ALPHA);
	if (m & (1u << 10)) BU_ADD(10, 0, KLASS_ALPHA);
	if (m & (1u << 11)) for (uint32_t v = 0; v < 4; v++) BU_ADD(11, v, KLASS_ALPHA);
	if (m & (1u << 12)) BU_ADD(12, 0, KLASS_ALPHA);
	if (m & (1u << 13)) for (uint32_t v = 0; v < 4; v++) BU_ADD(13, v, KLASS_ALPHA);
	if (m & (1u << 14)) BU_ADD(14, 0, KLASS_ALPHA);
#undef BU_ADD
	return n;
}

BU_FI bool slot_active(const slot_desc& s, block_class k, const level_opts& o)
{
	if (s.klass == KLASS_LA) return k.is_la != 0;
	if (s.klass == KLASS_RGB) return !k.has_alpha;
	return k.has_alpha || o.always_try_alpha;
}

// Partition ranking needed by the estimated-partition levels (uastc_enc.cpp:678-683, 828-858, 1000-1005, 1362-1406, 1635-1654).
BU_HD inline void rank_block(const bu_tables* T, const level_opts& o, block_class k, const uint32_t* px, block_ranks& out)
{
	for (int i = 0; i < 12; i++) out.r[i] = 0;
	if (!o.estimate_partition) return;
	const uint32_t m = o.mode_mask;
	if (!k.has_alpha)
	{
		if (m & (1u << 2)) rank_partitions(T, 2, 3, 3, px, 1, out.r + RANK_M2);
		if (m & (1u << 3)) rank_partitions(T, 3, 2, 3, px, 1, out.r + RANK_M3);
		if (m & (1u << 4)) rank_partitions(T, 2, 2, 3, px, 1, out.r + RANK_M4);
		if (m & (1u << 7)) rank_partitions(T, 7, 2, 3, px, 1, out.r + RANK_M7);
	}
	if ((k.has_alpha || o.always_try_alpha) && (m & (1u << 9)))
		rank_partitions(T, 2, 2, 4, px, 4, out.r + RANK_M9);
	if (k.is_la && (m & (1u << 16)))
	{
		uint32_t la[16];
		for (int i = 0; i < 16; i++) la[i] = px_make(px_c(px[i], 0), 0, 0, px_c(px[i], 3));
		rank_partitions(T, 2, 2, 4, la, 4, out.r + RANK_M16);
	}
}

// Generate + score one slot's candidate.
BU_HD inline void run_slot(const bu_tables* T, const level_opts& o, const slot_desc& s, block_class k, const block_ranks& ranks, const uint32_t* px, candidate& c)
{
	const uint32_t mode = s.mode;
	const bool est = o.estimate_partition != 0;
	switch (mode)
	{
	case 2: gen_multi_subset(T, mode, est ? ranks.r[RANK_M2] : s.variant, o, px, c); break;
	case 3: gen_multi_subset(T, mode, est ? ranks.r[RANK_M3] : s.variant, o, px, c); break;
	case 4: gen_multi_subset(T, mode, est ? ranks.r[RANK_M4] : s.variant, o, px, c); break;
	case 7: gen_multi_subset(T, mode, est ? ranks.r[RANK_M7] : s.variant, o, px, c); break;
	case 9: gen_multi_subset(T, mode, est ? ranks.r[RANK_M9 + s.variant] : s.variant, o, px, c); break;
	case 16: gen_multi_subset(T, mode, est ? ranks.r[RANK_M16 + s.variant] : s.variant, o, px, c); break;
	case 6: case 11: case 13: case 17: gen_dual_plane(T, mode, s.variant, o, px, c); break;
	default: gen_one_subset(T, mode, o, px, c); break;
	}
	score_candidate(T, px, k, c);
}

// Hints + packing for the chosen candidate (uastc_enc.cpp:3551-3639).
BU_HD inline void finish_block(const bu_tables* T, const level_opts& o, int level, uint32_t flags, const uint32_t* px, const candidate& best, uint8_t* out16)
{
	uint32_t dec[16];
	decode_candidate(T, best, dec);

	uint8_t ep[18], w[32];
	canonicalize(T, best, ep, w);

	bool h0 = false, h1 = false;
	if (o.bc1_hints) compute_bc1_hints(T, best.mode, ep, w, px, dec, h0, h1);

	uint32_t eac_table = 0, eac_mul = 0;
	if (T->mode_has_alpha[best.mode]) compute_eac_hint(T, dec, o.eac_mul_rad, o.eac_table_mask, eac_table, eac_mul);

	const etc1_search_opts so = etc1_search_setup(T, best.mode, level, flags, dec);
	ycc src_y[16], dec_y[16];
	for (int i = 0; i < 16; i++) { src_y[i] = to_ycc(px[i]); dec_y[i] = to_ycc(dec[i]); }
	etc1_hint hint;
	hint.err = UINT64_MAX; hint.order = 0; hint.flip = 0; hint.diff = 0; hint.inten0 = 0; hint.inten1 = 0; hint.bias = 0;
	for (uint32_t flip = so.first_flip; flip < so.last_flip; flip++)
		for (uint32_t individ = so.first_individ; individ < so.last_individ; individ++)
			etc1_hint_trials(T, so, flip, individ, 0, 1, src_y, dec_y, dec, hint);

	pack_block(T, best, ep, w, hint, eac_table, eac_mul, h0, h1, out16);
}

} // namespace bu
