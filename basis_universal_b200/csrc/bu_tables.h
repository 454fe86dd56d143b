// bu_tables.h -- constant codec tables shared by every kernel (one aggregate so it can live in global memory,
// be staged into shared memory wholesale, or sit in host memory for the host-emulation build).
// Field order MUST match oracle/gen_tables.cpp, which emits the initialiser (uastc_tables.inc) from the
// reference's own initialised tables.  Meaning of each field cites the reference table it mirrors.
#pragma once
#include <stdint.h>

struct bu_tables
{
	// ASTC endpoint quantisation for the 8 BISE ranges UASTC uses {7,8,11,12,13,18,19,20} (transcoder.cpp:14417).
	int8_t   range_slot[21];          // range -> slot, -1 if unused
	uint16_t range_levels[8];         // astc_get_levels (transcoder.cpp:14460)
	uint8_t  unq[8 * 256];            // [slot][ASTC index]    -> 8-bit value   (g_astc_unquant, transcoder.cpp:14467)
	alignas(16) uint8_t  sorted_unq[8 * 256];     // (16-byte aligned rows: bulk-copied to shared memory by k_candidates)
	                                  // [slot][sorted order]  -> 8-bit value   (g_astc_sorted_order_unquant, bc7enc.cpp:93)
	uint8_t  sorted_idx[8 * 256];     // [slot][sorted order]  -> ASTC index
	alignas(16) uint8_t  nearest[8 * 256];        // [slot][8-bit value]   -> sorted order  (g_astc_nearest_sorted_index, bc7enc.cpp:95)

	// Optimal single-colour endpoints {lo,hi} in sorted order (bc7enc.cpp:75-91, built at bc7enc.cpp:219-391).
	uint8_t  one_r8_w3[512], one_r8_w2[512], one_r7_w2[512], one_r13_w2[512], one_r11_w5[512];

	// Interpolation weights by weight-bit count 1..5 (transcoder.cpp:14582-14587) and the float LS forms (bc7enc.cpp:44-65).
	alignas(16) uint8_t  weights[6 * 32];
	alignas(16) float    weightsx[6 * 32 * 4];
	uint8_t  bc7_weights4[16];

	// UASTC mode properties (transcoder.cpp:14415-14427), Huffman mode codes (14380), BISE bits/trits/quints (14430).
	uint8_t  mode_weight_bits[19], mode_endpoint_range[19], mode_subsets[19], mode_planes[19], mode_comps[19];
	uint8_t  mode_has_etc1_bias[19], mode_has_bc1_hint0[19], mode_has_bc1_hint1[19], mode_has_alpha[19], mode_is_la[19];
	uint8_t  mode_huff[20 * 2];
	uint8_t  bise[21 * 3];

	// Common ASTC/BC7 partitions (transcoder.cpp:14271-14378). *_part* rows are texel->subset in BC7 numbering,
	// astc_pat* in ASTC numbering; anchors are the first texel of each ASTC subset.
	uint8_t  cp2_bc7[30], cp2_invert[30], cp3_bc7[11], cp3_perm[11], cp73_bc7[19], cp73_k[19];
	uint8_t  astc_to_bc7_perm[6 * 3];
	uint8_t  bc7_part2[30 * 16], bc7_part3[11 * 16], bc7_part73[19 * 16];
	uint8_t  astc_pat2[30 * 16], astc_pat3[11 * 16], astc_pat73[19 * 16];
	uint8_t  anchors2[30 * 3], anchors3[11 * 3], anchors73[19 * 3];

	// BC1 hint helpers: {hi,lo} optimal 5/6-bit endpoints for selector 1 (transcoder.cpp:1226), weight->BC1 selector (17730).
	uint8_t  bc1_match5[512], bc1_match6[512];
	uint8_t  uastc_to_bc1[6 * 32];

	// ETC1/EAC (etc.cpp:27, 304, 311).
	int16_t  etc1_inten[8 * 4];
	uint8_t  selector_index_to_etc1[4];
	int8_t   eac_tables[16 * 8];

	// ETC1 solid-colour packer (etc.cpp:40-118): per 8-bit value, offset of a 0xFFFF-terminated list of
	// diff|inten<<1|selector<<4|packed_c<<8 configurations; inverse lookup [diff|inten<<1|selector<<4][colour] -> packed_c | err<<8.
	uint16_t solid_cfg_ofs[256];
	uint16_t solid_cfg[2048];
	uint16_t etc1_inverse[64 * 256];

	// ETC1S optimiser: selector-count permutations tried by the cluster fit (etc.cpp:267) and the intensity-table pruning
	// table indexed by [inten_table][max component spread] (etc.cpp:1091).
	uint8_t  cluster_fit_order[165 * 4];
	uint8_t  eval_dist[8 * 256];
};
