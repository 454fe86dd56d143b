// bu_etc1s.h -- ETC1S frontend per-block stages: the arithmetic behind the reference's GPU seam (encoder/basisu_opencl.h:46-141).
//
// The contract of each `opencl_*` entry point is what its OpenCL C kernel computes (bin/ocl_kernels.cl); the functions here
// are bit-exact with those kernels (integer arithmetic throughout, except the block average: one u64->f32 divide and
// `(int)(avg * (31/255.0f) + .5f)` roundings, reproduced without FMA contraction):
//   etc1s_optimise            etc1s_optimizer_init / _internal_cluster_fit / _evaluate_solution   cl:772-982
//   refine / fosc / selectors the three integer kernels                                            cl:1063, 1159, 1227
// Early-outs in the kernels only drop partial sums that have already lost (`>=` against the running best), so they are
// omitted; argmin rules (first strictly smaller; the block's own cluster wins ties) are kept.
#pragma once
#include "bu_common.h"
#include "bu_tables.h"

namespace bu {

// basisu::color_distance (encoder/basisu_enc.h:1141), RGB only; identical to the .cl restatement (cl:71).
BU_FI uint32_t etc_color_distance(bool perceptual, uint32_t a, uint32_t b)
{
	const int dr = (int)px_c(a, 0) - (int)px_c(b, 0), dg = (int)px_c(a, 1) - (int)px_c(b, 1), db = (int)px_c(a, 2) - (int)px_c(b, 2);
	if (perceptual)
	{
		const int dl = dr * 14 + dg * 45 + db * 5;
		const int dcr = dr * 64 - dl, dcb = db * 64 - dl;
		return ((uint32_t)(dl * dl) >> 5) + ((((uint32_t)(dcr * dcr) >> 5) * 26u) >> 7) + ((((uint32_t)(dcb * dcb) >> 5) * 3u) >> 7);
	}
	return (uint32_t)(dr * dr + dg * dg + db * db);
}

// get_block_colors5 (cl:510): 5-bit base -> four 8-bit block colours of intensity table `inten`.
BU_FI void etc1s_block_colors(const bu_tables* T, uint32_t r5, uint32_t g5, uint32_t b5, uint32_t inten, uint32_t* colors)
{
	const int r = (int)((r5 << 3) | (r5 >> 2)), g = (int)((g5 << 3) | (g5 >> 2)), b = (int)((b5 << 3) | (b5 >> 2));
	for (int s = 0; s < 4; s++)
	{
		const int m = T->etc1_inten[inten * 4 + s];
		colors[s] = px_make(clamp255i(r + m), clamp255i(g + m), clamp255i(b + m), 255);
	}
}

// ETC1S block bytes (etc.h:91 big-endian bitfield): flip = 1, diff = 1, delta = 0, both subblocks share `inten`.
// raw_sel[i] are RAW ETC1 selector codes (0..3) for texel i = x + 4y (etc_block_pack_raw_selectors, cl:700).
BU_FI uint64_t etc1s_pack(uint32_t r5, uint32_t g5, uint32_t b5, uint32_t inten, uint32_t lsb_plane, uint32_t msb_plane)
{
	// returned as the 8 bytes in memory order packed little-endian into a u64
	const uint64_t b0 = r5 << 3, b1 = g5 << 3, b2 = b5 << 3, b3 = (inten << 5) | (inten << 2) | 3;
	const uint64_t b4 = (msb_plane >> 8) & 255, b5_ = msb_plane & 255, b6 = (lsb_plane >> 8) & 255, b7 = lsb_plane & 255;
	return b0 | (b1 << 8) | (b2 << 16) | (b3 << 24) | (b4 << 32) | (b5_ << 40) | (b6 << 48) | (b7 << 56);
}

// Two arithmetic "flavours" of the block/cluster optimiser exist in the reference and the seam must be able to stand in for
// either:
//   ETC1S_FLAVOUR_OCL  what the OpenCL kernels compute (cl:772-982): intensity tables always pruned by g_eval_dist_tables,
//                      base colours rounded with x * (31/255.0f).
//   ETC1S_FLAVOUR_CPU  what the CPU etc1_optimizer computes for the same stage (etc.cpp:948-995, 1104-1278) at quality
//                      Medium/Slow/Uber: pruning only at Medium (total_perms <= 16), rounding x * 31 / 255.0f, and the
//                      2-probe Bloom filter over tried base colours (etc.cpp:1072-1089, hash_hsieh of the 3 colour bytes)
//                      whose false positives skip evaluations. Bit-exact with the CPU frontend => ETC1S output (and PSNR)
//                      identical to the reference CPU encoder's for these stages. Quality Fast (total_perms == 4) uses a
//                      different sorted-luma evaluator on the CPU; it keeps the OCL flavour.
enum { ETC1S_FLAVOUR_OCL = 0, ETC1S_FLAVOUR_CPU = 1 };

// basist::hash_hsieh (transcoder/basisu_transcoder.cpp:355) specialised to the 3-byte key (r5, g5, b5).
BU_FI uint32_t hash_hsieh3(uint32_t r5, uint32_t g5, uint32_t b5)
{
	uint32_t h = 3;
	h += r5 | (g5 << 8);
	h ^= h << 16;
	h ^= b5 << 18;
	h += h >> 11;
	h ^= h << 3; h += h >> 5; h ^= h << 4; h += h >> 17; h ^= h << 25; h += h >> 6;
	return h;
}

struct etc1s_bloom { uint32_t bits[32]; };
BU_FI void bloom_clear(etc1s_bloom& b) { for (int i = 0; i < 32; i++) b.bits[i] = 0; }
// false if (r5,g5,b5) was probably tried already (etc.cpp:1072 check_for_redundant_solution)
BU_FI bool bloom_test_and_set(etc1s_bloom& b, uint32_t r5, uint32_t g5, uint32_t b5)
{
	const uint32_t kh = hash_hsieh3(r5, g5, b5);
	const uint32_t h0 = kh & 1023, h1 = (kh >> 10) & 1023;
	if ((b.bits[h0 >> 5] >> (h0 & 31)) & (b.bits[h1 >> 5] >> (h1 & 31)) & 1) return false;
	b.bits[h0 >> 5] |= 1u << (h0 & 31);
	b.bits[h1 >> 5] |= 1u << (h1 & 31);
	return true;
}

BU_FI uint32_t etc1s_round5(float v, int flavour)
{
	const int i = (flavour == ETC1S_FLAVOUR_CPU) ? (int)(v * 31.0f / 255.0f + .5f) : (int)(v * (31 / 255.0f) + .5f);
	return (uint32_t)clampi(i, 0, 31);
}

struct etc1s_solution
{
	uint64_t err;
	uint32_t r5, g5, b5, inten;
	uint32_t lsb_plane, msb_plane; // raw selector bit planes, bit index = x*4+y (only maintained for 16-texel blocks)
	uint32_t valid;
};

// One candidate base colour against all (non-pruned) intensity tables, 16-texel block (cl:772). Updates `best` if strictly better.
BU_HD inline void etc1s_evaluate_block(const bu_tables* T, bool perceptual, const uint32_t* px, uint32_t spread, uint32_t r5, uint32_t g5, uint32_t b5, etc1s_solution& best, bool prune = true)
{
	uint64_t trial_err = (uint64_t)INT64_MAX;
	uint32_t trial_inten = 0, trial_lsb = 0, trial_msb = 0;
	bool valid = false;
	for (uint32_t inten = 0; inten < 8; inten++)
	{
		if (prune && !T->eval_dist[inten * 256 + spread]) continue;
		uint32_t colors[4];
		etc1s_block_colors(T, r5, g5, b5, inten, colors);
		uint64_t total = 0;
		uint32_t lsb = 0, msb = 0;
		for (int i = 0; i < 16; i++)
		{
			// selector index s -> raw ETC1 code {3,2,0,1}[s]; first strictly smaller error wins (cl:800-825)
			uint32_t be = etc_color_distance(perceptual, px[i], colors[0]), raw = 3;
			uint32_t e = etc_color_distance(perceptual, px[i], colors[1]); if (e < be) { be = e; raw = 2; }
			e = etc_color_distance(perceptual, px[i], colors[2]); if (e < be) { be = e; raw = 0; }
			e = etc_color_distance(perceptual, px[i], colors[3]); if (e < be) { be = e; raw = 1; }
			total += be;
			const uint32_t bit = (uint32_t)((i & 3) * 4 + (i >> 2));
			lsb |= (raw & 1) << bit; msb |= (raw >> 1) << bit;
		}
		if (total < trial_err) { trial_err = total; trial_inten = inten; trial_lsb = lsb; trial_msb = msb; valid = true; }
	}
	if (trial_err < best.err)
	{
		best.err = trial_err; best.r5 = r5; best.g5 = g5; best.b5 = b5; best.inten = trial_inten;
		best.lsb_plane = trial_lsb; best.msb_plane = trial_msb; best.valid = valid ? 1u : 0u;
	}
}

// Next base colour proposed by the cluster fit for selector-count permutation `perm` (cl:944-973). Returns false if the
// permutation leaves the colour unchanged (all clamped deltas zero).
BU_FI bool etc1s_cluster_fit_step(const bu_tables* T, const etc1s_solution& best, const float* avg, uint32_t perm, uint32_t& r1, uint32_t& g1, uint32_t& b1, int flavour = ETC1S_FLAVOUR_OCL)
{
	const int br = (int)((best.r5 << 3) | (best.r5 >> 2)), bg = (int)((best.g5 << 3) | (best.g5 >> 2)), bb = (int)((best.b5 << 3) | (best.b5 >> 2));
	int dr = 0, dg = 0, db = 0;
	for (int q = 0; q < 4; q++)
	{
		const int yd = T->etc1_inten[best.inten * 4 + q];
		const int cnt = T->cluster_fit_order[perm * 4 + q];
		dr += cnt * (clampi(br + yd, 0, 255) - br);
		dg += cnt * (clampi(bg + yd, 0, 255) - bg);
		db += cnt * (clampi(bb + yd, 0, 255) - bb);
	}
	if (!dr && !dg && !db) return false;
	r1 = etc1s_round5(avg[0] - (float)dr / 8, flavour);
	g1 = etc1s_round5(avg[1] - (float)dg / 8, flavour);
	b1 = etc1s_round5(avg[2] - (float)db / 8, flavour);
	return true;
}

// encode_etc1s_blocks (cl:984) / init_etc1_images CPU path (frontend.cpp:775-815): optimise one 4x4 block, return the 8 block bytes.
BU_HD inline uint64_t etc1s_encode_block(const bu_tables* T, bool perceptual, uint32_t total_perms, const uint32_t* px, int flavour = ETC1S_FLAVOUR_OCL)
{
	if (total_perms <= 4) flavour = ETC1S_FLAVOUR_OCL; // CPU quality Fast is a different evaluator (see the flavour note)
	uint32_t mn[3] = { 255, 255, 255 }, mx[3] = { 0, 0, 0 };
	uint64_t sum[3] = { 0, 0, 0 };
	for (int i = 0; i < 16; i++)
		for (uint32_t c = 0; c < 3; c++)
		{
			const uint32_t v = px_c(px[i], c);
			mn[c] = minu(mn[c], v); mx[c] = maxu(mx[c], v); sum[c] += v;
		}
	float avg[3];
	for (int c = 0; c < 3; c++) avg[c] = (float)sum[c] / (float)(uint64_t)16; // 16-texel sums are exact in float either way
	const uint32_t spread = (uint32_t)maxi(maxi((int)mx[0] - (int)mn[0], (int)mx[1] - (int)mn[1]), (int)mx[2] - (int)mn[2]);
	const uint32_t r0 = etc1s_round5(avg[0], flavour), g0 = etc1s_round5(avg[1], flavour), b0 = etc1s_round5(avg[2], flavour);
	const bool cpu = flavour == ETC1S_FLAVOUR_CPU;
	const bool prune = !cpu || total_perms <= 16;

	etc1s_bloom bloom;
	if (cpu) { bloom_clear(bloom); bloom_test_and_set(bloom, r0, g0, b0); }
	etc1s_solution best;
	best.err = UINT64_MAX; best.r5 = best.g5 = best.b5 = best.inten = 0; best.lsb_plane = best.msb_plane = 0; best.valid = 0;
	etc1s_evaluate_block(T, perceptual, px, spread, r0, g0, b0, best, prune);
	if (best.err != 0)
		for (uint32_t perm = 0; perm < total_perms; perm++)
		{
			uint32_t r1, g1, b1;
			if (!etc1s_cluster_fit_step(T, best, avg, perm, r1, g1, b1, flavour)) continue;
			if (cpu && !bloom_test_and_set(bloom, r1, g1, b1)) continue;
			etc1s_evaluate_block(T, perceptual, px, spread, r1, g1, b1, best, prune);
			if (best.err == 0) break;
		}
	return etc1s_pack(best.r5, best.g5, best.b5, best.inten, best.lsb_plane, best.msb_plane);
}

// determine_selectors (cl:1227): selectors for a given (rgb5, inten); ties prefer the lowest selector index.
BU_HD inline uint64_t etc1s_determine_selectors(const bu_tables* T, bool perceptual, const uint32_t* px, uint32_t color5_inten)
{
	const uint32_t r5 = px_c(color5_inten, 0), g5 = px_c(color5_inten, 1), b5 = px_c(color5_inten, 2), inten = px_c(color5_inten, 3);
	uint32_t colors[4];
	etc1s_block_colors(T, r5, g5, b5, inten, colors);
	uint32_t lsb = 0, msb = 0;
	for (int i = 0; i < 16; i++)
	{
		uint32_t be = etc_color_distance(perceptual, px[i], colors[0]), s = 0;
		for (uint32_t k = 1; k < 4; k++) { const uint32_t e = etc_color_distance(perceptual, px[i], colors[k]); if (e < be) { be = e; s = k; } }
		const uint32_t raw = T->selector_index_to_etc1[s];
		const uint32_t bit = (uint32_t)((i & 3) * 4 + (i >> 2));
		lsb |= (raw & 1) << bit; msb |= (raw >> 1) << bit;
	}
	// pack_color5(c_unscaled, scaled=false) clamps each component to 31 (etc.cpp:368)
	return etc1s_pack(minu(r5, 31), minu(g5, 31), minu(b5, 31), inten & 7, lsb, msb);
}

// Sum over the 16 texels of the best-of-4 error for one (rgb5, inten) (inner loop of refine_endpoint_clusterization, cl:1103-1126).
BU_FI uint64_t etc1s_block_error(const bu_tables* T, bool perceptual, const uint32_t* px, uint32_t r5, uint32_t g5, uint32_t b5, uint32_t inten)
{
	uint32_t colors[4];
	etc1s_block_colors(T, r5, g5, b5, inten, colors);
	uint64_t total = 0;
	for (int i = 0; i < 16; i++)
	{
		uint32_t be = etc_color_distance(perceptual, px[i], colors[0]);
		be = minu(be, etc_color_distance(perceptual, px[i], colors[1]));
		be = minu(be, etc_color_distance(perceptual, px[i], colors[2]));
		be = minu(be, etc_color_distance(perceptual, px[i], colors[3]));
		total += be;
	}
	return total;
}

// ---- cluster optimiser (etc1_optimizer over an arbitrary texel set) ---------------------------------------------------------
// A cluster is optimised by a TEAM whose members stride its texels (b200_etc1s.cu: a warp or a CTA; the host emulation: one
// thread). Src supplies pixel(i) / weight(i) / selector(i) and the traits `forced` (selectors imposed) and `unit_weights`.
// Weighted error of base colour (r5,g5,b5) against the cluster's texels for every non-pruned intensity table; returns the
// best (first strictly smaller) table and its error. Uniform across the team.
template<typename Src, typename Team> BU_HD inline void cluster_evaluate(const bu_tables* T, bool perceptual, const Src& src, Team& team,
	uint32_t spread, uint32_t r5, uint32_t g5, uint32_t b5, etc1s_solution& best, bool prune)
{
	const uint64_t n = src.n;
	uint64_t trial_err = (uint64_t)INT64_MAX;
	uint32_t trial_inten = 0;
	bool valid = false;
	for (uint32_t inten = 0; inten < 8; inten++)
	{
		if (prune && !T->eval_dist[inten * 256 + spread]) continue;
		uint32_t colors[4];
		etc1s_block_colors(T, r5, g5, b5, inten, colors);
		uint64_t total = 0;
		for (uint64_t i = team.rank; i < n; i += Team::size)
		{
			const uint32_t p = src.pixel(i);
			uint32_t be;
			if (Src::forced)
			{
				const uint32_t sl = src.selector(i);
				be = etc_color_distance(perceptual, p, sl == 0 ? colors[0] : (sl == 1 ? colors[1] : (sl == 2 ? colors[2] : colors[3])));
			}
			else
			{
				be = etc_color_distance(perceptual, p, colors[0]);
				be = minu(be, etc_color_distance(perceptual, p, colors[1]));
				be = minu(be, etc_color_distance(perceptual, p, colors[2]));
				be = minu(be, etc_color_distance(perceptual, p, colors[3]));
			}
			total += (uint64_t)be * (uint64_t)src.weight(i);
		}
		total = team.sum(total);
		if (total < trial_err) { trial_err = total; trial_inten = inten; valid = true; }
	}
	if (trial_err < best.err) { best.err = trial_err; best.r5 = r5; best.g5 = g5; best.b5 = b5; best.inten = trial_inten; best.valid = valid ? 1u : 0u; }
}

// One team optimises one cluster (etc1_optimizer over the cluster's texels); returns the packed base colour + intensity table.
template<typename Src, typename Team> BU_HD inline uint64_t cluster_optimize(const bu_tables* T, bool perceptual, const Src& src, Team& team, uint32_t total_perms, int flavour, uint64_t* pErr = nullptr)
{
	const uint64_t n = src.n;
	uint32_t mn[3] = { 255, 255, 255 }, mx[3] = { 0, 0, 0 };
	uint64_t sum[3] = { 0, 0, 0 }, tw = 0;
	for (uint64_t i = team.rank; i < n; i += Team::size)
	{
		const uint32_t p = src.pixel(i);
		const uint64_t w = src.weight(i);
		for (uint32_t c = 0; c < 3; c++)
		{
			const uint32_t v = px_c(p, c);
			mn[c] = minu(mn[c], v); mx[c] = maxu(mx[c], v); sum[c] += w * v;
		}
		tw += w;
	}
	for (int c = 0; c < 3; c++)
	{
		team.minmax(mn[c], mx[c]);
		sum[c] = team.sum(sum[c]);
	}
	tw = team.sum(tw);

	float avg[3];
	for (int c = 0; c < 3; c++) avg[c] = (float)sum[c] / (float)tw;
	if (Src::unit_weights && flavour == ETC1S_FLAVOUR_CPU && (sum[0] > (1ull << 24) || sum[1] > (1ull << 24) || sum[2] > (1ull << 24)))
	{
		// etc1_optimizer::init (etc.cpp:1022-1040) adds the texels to a float vector in member order; past 2^24 that sum rounds, so
		// such a cluster (> 4112 blocks) repeats the serial float additions - every team member redundantly, the loads are broadcasts.
		float f0 = 0.0f, f1 = 0.0f, f2 = 0.0f;
		for (uint64_t i = 0; i < n; i++)
		{
			const uint32_t p = src.pixel(i);
			f0 += (float)px_c(p, 0); f1 += (float)px_c(p, 1); f2 += (float)px_c(p, 2);
		}
		avg[0] = f0 / (float)n; avg[1] = f1 / (float)n; avg[2] = f2 / (float)n;
	}
	const uint32_t spread = maxu(maxu(mx[0] - mn[0], mx[1] - mn[1]), mx[2] - mn[2]);
	if (total_perms <= 4) flavour = ETC1S_FLAVOUR_OCL;
	const bool cpu = flavour == ETC1S_FLAVOUR_CPU;
	const bool prune = !cpu || total_perms <= 16;
	const uint32_t r0 = etc1s_round5(avg[0], flavour), g0 = etc1s_round5(avg[1], flavour), b0 = etc1s_round5(avg[2], flavour);

	if (cpu) team.bloom_test_and_set(r0, g0, b0);
	etc1s_solution best;
	best.err = UINT64_MAX; best.r5 = best.g5 = best.b5 = best.inten = 0; best.lsb_plane = best.msb_plane = 0; best.valid = 0;
	cluster_evaluate(T, perceptual, src, team, spread, r0, g0, b0, best, prune);
	if (best.err != 0)
		for (uint32_t perm = 0; perm < total_perms; perm++)
		{
			uint32_t r1, g1, b1;
			if (!etc1s_cluster_fit_step(T, best, avg, perm, r1, g1, b1, flavour)) continue;
			if (cpu && !team.bloom_test_and_set(r1, g1, b1)) continue;
			cluster_evaluate(T, perceptual, src, team, spread, r1, g1, b1, best, prune);
			if (best.err == 0) break;
		}
	if (pErr) *pErr = best.err;
	return etc1s_pack(best.r5, best.g5, best.b5, best.inten, 0, 0); // the kernel defines no selectors for clusters
}

// Error of one given (colour, table) over the cluster's texels: with the imposed selectors for a forced source (the "current error"
// of reoptimize_remapped_endpoints, frontend.cpp:3033-3046), else with the best of the four colours per texel (the "previous error"
// of generate_endpoint_codebook at step >= 1, frontend.cpp:1560-1590).
template<typename Src, typename Team> BU_HD inline uint64_t cluster_endpoint_error(const bu_tables* T, bool perceptual, const Src& src, Team& team, uint32_t r5, uint32_t g5, uint32_t b5, uint32_t inten)
{
	uint32_t colors[4];
	etc1s_block_colors(T, r5, g5, b5, inten, colors);
	uint64_t total = 0;
	for (uint64_t i = team.rank; i < src.n; i += Team::size)
	{
		const uint32_t p = src.pixel(i);
		uint32_t be;
		if (Src::forced)
		{
			const uint32_t sl = src.selector(i);
			be = etc_color_distance(perceptual, p, sl == 0 ? colors[0] : (sl == 1 ? colors[1] : (sl == 2 ? colors[2] : colors[3])));
		}
		else
		{
			be = etc_color_distance(perceptual, p, colors[0]);
			be = minu(be, etc_color_distance(perceptual, p, colors[1]));
			be = minu(be, etc_color_distance(perceptual, p, colors[2]));
			be = minu(be, etc_color_distance(perceptual, p, colors[3]));
		}
		total += be;
	}
	return team.sum(total);
}

// ---- ETC1S backend, endpoint prediction (basisu_backend::create_encoder_blocks, backend.cpp:437-600): the per-block decision ----------
// The kernel (b200_etc1s.cu) supplies the neighbours' already-decided indices; the host emulation runs the same functions in raster order.

// Selector INDEX (0..3, darkest..brightest) of texel (x, y) of an etc_block given as two little-endian words: bytes 4..7 = msb plane
// (hi, lo), lsb plane (hi, lo), bit x * 4 + y; raw code -> selector index {2, 3, 1, 0} (etc.h:91, g_etc1_to_selector_index).
BU_FI uint32_t etc1s_texel_selector(uint32_t etc_hi, uint32_t x, uint32_t y)
{
	const uint32_t msb = ((etc_hi & 255u) << 8) | ((etc_hi >> 8) & 255u), lsb = (((etc_hi >> 16) & 255u) << 8) | (etc_hi >> 24);
	const uint32_t bit = x * 4 + y;
	const uint32_t raw = ((lsb >> bit) & 1u) | (((msb >> bit) & 1u) << 1);
	return (0x1Eu >> (raw * 2)) & 3u; // raw 0 -> 2, 1 -> 3, 2 -> 1, 3 -> 0
}

// Error of a block decoded with endpoint (r5, g5, b5, inten) and the block's own selectors (etc_block::evaluate_etc1_error /
// the unpack_etc1 + color_distance loop of backend.cpp:540-563). Sums row by row and stops once past `give_up`: the reference stops
// per texel, and what it then compares is already too large either way.
BU_HD inline uint64_t etc1s_block_error_with_selectors(const bu_tables* T, bool perceptual, const uint32_t* px, uint32_t etc_hi, uint32_t r5, uint32_t g5, uint32_t b5, uint32_t inten, uint64_t give_up)
{
	uint32_t colors[4];
	etc1s_block_colors(T, r5, g5, b5, inten, colors);
	uint64_t total = 0;
	for (uint32_t y = 0; y < 4; y++)
	{
		for (uint32_t x = 0; x < 4; x++)
		{
			const uint32_t s = etc1s_texel_selector(etc_hi, x, y);
			total += etc_color_distance(perceptual, px[x + y * 4], s == 0 ? colors[0] : (s == 1 ? colors[1] : (s == 2 ? colors[2] : colors[3])));
		}
		if (total > give_up) break;
	}
	return total;
}

// Predictor 0 / 1 / 2 = left / upper / upper-left neighbour with the block's own endpoint index, lowest first; 3 = none (backend.cpp:453-495).
BU_FI uint32_t etc1s_predict_from_neighbours(uint32_t own, const uint32_t* nb, const bool* has)
{
	if (has[0] && nb[0] == own) return 0;
	if (has[1] && nb[1] == own) return 1;
	if (has[2] && nb[2] == own) return 2;
	return 3;
}

// Endpoint RDO of a block no neighbour predicts (backend.cpp:497-590): the neighbour endpoint with the lowest error within
// max(1, thresh) x the block's current error, lowest predictor on ties. Returns the predictor (new_index set), 3 if none qualifies,
// 3 | 0x80 if the block's current error is zero (the reference then neither searches nor counts the block).
// etc_lo / etc_hi: the block's 8 bytes as two little-endian words (bytes 0..3 = R5 << 3, G5 << 3, B5 << 3 with zero delta, table bits).
BU_HD inline uint32_t etc1s_endpoint_rdo(const bu_tables* T, bool perceptual, const uint32_t* px, uint32_t etc_lo, uint32_t etc_hi, const uint32_t* nb, const uint32_t* nb_c5i, const bool* has,
	float thresh, uint32_t& new_index)
{
	const uint64_t cur_err = etc1s_block_error_with_selectors(T, perceptual, px, etc_hi, (etc_lo >> 3) & 31u, (etc_lo >> 11) & 31u, (etc_lo >> 19) & 31u, (etc_lo >> 29) & 7u, UINT64_MAX);
	if (!cur_err) return 3u | 0x80u;
	const uint64_t thresh_err = (uint64_t)((float)cur_err * (thresh > 1.0f ? thresh : 1.0f)); // uint64 -> float, float product, truncation: as written at backend.cpp:509
	uint64_t best_err = UINT64_MAX;
	uint32_t pred = 3;
	for (uint32_t p = 0; p < 3; p++)
	{
		if (!has[p]) continue;
		const uint32_t e = nb_c5i[p];
		const uint64_t trial = etc1s_block_error_with_selectors(T, perceptual, px, etc_hi, e & 255u, (e >> 8) & 255u, (e >> 16) & 255u, e >> 24, thresh_err);
		if (trial <= thresh_err && trial < best_err) { best_err = trial; new_index = nb[p]; pred = p; }
	}
	return pred;
}

} // namespace bu
