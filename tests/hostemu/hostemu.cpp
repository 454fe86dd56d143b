// tests/hostemu/hostemu.cpp -- TEST INFRASTRUCTURE.  Compiles the *device* headers (basis_universal_b200/csrc/bu_*.h) for
// the host with g++ (-ffp-contract=off, baseline x86-64: same IEEE semantics the kernels are built with) so the CPU-only
// test suite can check the algorithm against the compiled reference without a GPU, block by block and stage by stage.
// It is never loaded by the product package; the product path is the CUDA library only.
#include "../../basis_universal_b200/csrc/bu_slots.h"
#include <vector>
#include <thread>
#include <atomic>
#include <cstdio>

static const bu_tables g_tables =
#include "../../basis_universal_b200/csrc/uastc_tables.inc"
;

#define EMU_API extern "C" __attribute__((visibility("default")))

using namespace bu;

static void encode_one(const uint32_t* px, uint8_t* out16, uint32_t flags, const level_opts& o, int level, const slot_desc* slots, uint32_t nslots, int* dbg_best)
{
	const bu_tables* T = &g_tables;
	const block_class k = classify_block(px, o.la_only_transparent != 0);
	if (k.solid) { pack_solid_block(T, px[0], out16); if (dbg_best) *dbg_best = -1; return; }

	block_ranks ranks;
	rank_block(T, o, k, px, ranks);

	std::vector<candidate> cands;
	cands.reserve(nslots);
	for (uint32_t i = 0; i < nslots; i++)
	{
		if (!slot_active(slots[i], k, o)) continue;
		candidate c;
		run_slot(T, o, slots[i], k, ranks, px, c);
		cands.push_back(c);
	}
	std::vector<uint32_t> ue(cands.size()), be(cands.size());
	std::vector<uint8_t> modes(cands.size());
	for (size_t i = 0; i < cands.size(); i++) { ue[i] = cands[i].uastc_err; be[i] = cands[i].bc7_err; modes[i] = cands[i].mode; }
	const int best = select_candidate((uint32_t)cands.size(), ue.data(), be.data(), modes.data(), flags);
	if (dbg_best) *dbg_best = best;
	finish_block(T, o, level, flags, px, cands[best], out16);
}

EMU_API void emu_encode_uastc_blocks(const uint8_t* pBlocks, uint32_t n, uint8_t* pOut, uint32_t flags, uint32_t threads)
{
	const int level = clampi((int)(flags & 7), 0, 4);
	const level_opts o = make_level_opts(level);
	slot_desc slots[MAX_SLOTS];
	const uint32_t nslots = build_slots(o, slots);
	std::atomic<uint32_t> next(0);
	auto work = [&]() {
		for (;;)
		{
			const uint32_t first = next.fetch_add(64);
			if (first >= n) break;
			const uint32_t last = first + 64 < n ? first + 64 : n;
			for (uint32_t i = first; i < last; i++)
			{
				uint32_t px[16];
				memcpy(px, pBlocks + (size_t)i * 64, 64);
				encode_one(px, pOut + (size_t)i * 16, flags, o, level, slots, nslots, nullptr);
			}
		}
	};
	if (threads <= 1) { work(); return; }
	std::vector<std::thread> pool;
	for (uint32_t t = 0; t < threads; t++) pool.emplace_back(work);
	for (auto& th : pool) th.join();
}

// Per-function hooks mirroring oracle/ref_shim.cpp for differential tests.
EMU_API uint64_t emu_color_cell_compression(const uint8_t* pPixels, uint32_t num_pixels, uint32_t weight_table, uint32_t endpoint_range, uint32_t has_alpha,
	uint32_t uber_level, uint32_t ls_passes, uint8_t* pLow4, uint8_t* pHigh4, uint8_t* pSelectors)
{
	cell_cfg cfg;
	cfg.wbits = weight_table; cfg.range = endpoint_range; cfg.slot = (uint32_t)g_tables.range_slot[endpoint_range];
	cfg.has_alpha = has_alpha; cfg.uber = uber_level; cfg.ls_passes = ls_passes;
	uint32_t px[16];
	memcpy(px, pPixels, num_pixels * 4);
	cell_result r;
	cell_compress(&g_tables, cfg, px, num_pixels, r);
	memcpy(pLow4, r.astc_lo, 4); memcpy(pHigh4, r.astc_hi, 4); memcpy(pSelectors, r.sel, num_pixels);
	return r.err;
}

EMU_API uint32_t emu_sizeof_candidate() { return (uint32_t)sizeof(candidate); }
EMU_API uint32_t emu_sizeof_tables() { return (uint32_t)sizeof(bu_tables); }

// ---- ETC1S stages (bu_etc1s.h) ----------------------------------------------------------------------------------------------
#include "../../basis_universal_b200/csrc/bu_etc1s.h"
#include "../../include/basisu_b200.h"

EMU_API void emu_etc1s_encode_blocks_flavour(const uint8_t* pBlocks, uint32_t n, uint8_t* pOut, int perceptual, uint32_t total_perms, int flavour)
{
	for (uint32_t i = 0; i < n; i++)
	{
		uint32_t px[16];
		memcpy(px, pBlocks + (size_t)i * 64, 64);
		const uint64_t v = etc1s_encode_block(&g_tables, perceptual != 0, total_perms, px, flavour);
		memcpy(pOut + (size_t)i * 8, &v, 8);
	}
}

EMU_API void emu_etc1s_encode_blocks(const uint8_t* pBlocks, uint32_t n, uint8_t* pOut, int perceptual, uint32_t total_perms)
{
	for (uint32_t i = 0; i < n; i++)
	{
		uint32_t px[16];
		memcpy(px, pBlocks + (size_t)i * 64, 64);
		const uint64_t v = etc1s_encode_block(&g_tables, perceptual != 0, total_perms, px);
		memcpy(pOut + (size_t)i * 8, &v, 8);
	}
}

// The cluster optimiser templates of bu_etc1s.h run by a team of one (the device runs them with a warp or a CTA).
struct serial_team
{
	uint32_t rank = 0; etc1s_bloom bloom;
	static constexpr uint32_t size = 1;
	serial_team() { bloom_clear(bloom); }
	uint64_t sum(uint64_t v) { return v; }
	void minmax(uint32_t&, uint32_t&) {}
	bool bloom_test_and_set(uint32_t r5, uint32_t g5, uint32_t b5) { return bu::bloom_test_and_set(bloom, r5, g5, b5); }
};
struct host_blocks_src
{
	static constexpr bool forced = false, unit_weights = true;
	const uint32_t* px; uint64_t n;
	uint32_t pixel(uint64_t i) const { return px[i]; }
	uint32_t weight(uint64_t) const { return 1u; }
	uint32_t selector(uint64_t) const { return 0; }
};
struct host_forced_src
{
	static constexpr bool forced = true, unit_weights = true;
	const uint32_t* px; const uint32_t* sels; uint64_t n;
	uint32_t pixel(uint64_t i) const { return px[i]; }
	uint32_t weight(uint64_t) const { return 1u; }
	uint32_t selector(uint64_t i) const { return (sels[i >> 4] >> (2 * (uint32_t)(i & 15))) & 3u; }
};
// One endpoint cluster of `nblocks` source blocks. pSelectors == NULL: b200_etc1s_encode_endpoint_clusters' per-cluster work;
// else b200_etc1s_reoptimize_endpoint_clusters' (selectors imposed); *pCur_err = error of the endpoint cur4 (with the imposed
// selectors, else best of four per texel). out4 = {r5, g5, b5, inten}.
EMU_API uint64_t emu_etc1s_optimize_cluster(const uint8_t* pBlocks, uint32_t nblocks, const uint32_t* pSelectors, const uint8_t* cur4, int perceptual, uint32_t total_perms, int flavour,
	uint8_t* out4, uint64_t* pCur_err)
{
	serial_team team;
	uint64_t err = 0, packed;
	if (pSelectors)
	{
		host_forced_src src; src.px = reinterpret_cast<const uint32_t*>(pBlocks); src.sels = pSelectors; src.n = (uint64_t)nblocks * 16;
		if (pCur_err) *pCur_err = cluster_endpoint_error(&g_tables, perceptual != 0, src, team, cur4[0], cur4[1], cur4[2], cur4[3]);
		packed = cluster_optimize(&g_tables, perceptual != 0, src, team, total_perms, flavour, &err);
	}
	else
	{
		host_blocks_src src; src.px = reinterpret_cast<const uint32_t*>(pBlocks); src.n = (uint64_t)nblocks * 16;
		if (pCur_err && cur4) *pCur_err = cluster_endpoint_error(&g_tables, perceptual != 0, src, team, cur4[0], cur4[1], cur4[2], cur4[3]);
		packed = cluster_optimize(&g_tables, perceptual != 0, src, team, total_perms, flavour, &err);
	}
	const uint32_t w = (uint32_t)packed;
	out4[0] = (uint8_t)((w >> 3) & 31); out4[1] = (uint8_t)((w >> 11) & 31); out4[2] = (uint8_t)((w >> 19) & 31); out4[3] = (uint8_t)((w >> 29) & 7);
	return err;
}

// basisu_backend::create_encoder_blocks' endpoint prediction for one slice, in the reference's raster order, with the per-block
// decision functions the wavefront kernel runs (bu_etc1s.h). pIdx in/out; pPred: 0..2, 3 = none, 0x83 = none and zero current error.
EMU_API void emu_backend_endpoint_prediction(const uint8_t* pBlocks, const uint8_t* pEtc, uint32_t nbx, uint32_t nby, const uint8_t* pC5i, float thresh, int perceptual, uint32_t* pIdx, uint8_t* pPred)
{
	const uint32_t* c5i = reinterpret_cast<const uint32_t*>(pC5i);
	for (uint32_t y = 0; y < nby; y++)
		for (uint32_t x = 0; x < nbx; x++)
		{
			const uint32_t b = x + y * nbx, own = pIdx[b];
			const bool has[3] = { x > 0, y > 0, x > 0 && y > 0 };
			const uint32_t nb[3] = { has[0] ? pIdx[b - 1] : 0xFFFFFFFFu, has[1] ? pIdx[b - nbx] : 0xFFFFFFFFu, has[2] ? pIdx[b - nbx - 1] : 0xFFFFFFFFu };
			uint32_t pred = etc1s_predict_from_neighbours(own, nb, has);
			if (pred == 3 && thresh > 0.0f)
			{
				const uint32_t* px = reinterpret_cast<const uint32_t*>(pBlocks + (size_t)b * 64);
				const uint32_t* etc = reinterpret_cast<const uint32_t*>(pEtc + (size_t)b * 8);
				uint32_t nb_c5i[3];
				for (uint32_t p = 0; p < 3; p++) nb_c5i[p] = has[p] ? c5i[nb[p]] : 0u;
				uint32_t new_index = own;
				pred = etc1s_endpoint_rdo(&g_tables, perceptual != 0, px, etc[0], etc[1], nb, nb_c5i, has, thresh, new_index);
				if ((pred & 3u) != 3u) pIdx[b] = new_index;
			}
			pPred[b] = (uint8_t)pred;
		}
}

// The sparse formulation of palette_index_reorderer::init that b200_backend.cu's k_pal_order walks (adjacency lists instead of the
// dense table; every entry's placed neighbours kept in placed-list order by prepending / appending; the side decision as a float sum
// of int products over that list), run serially: checks the formulation itself against the reference on the CPU.
#include <algorithm>
#include <vector>
EMU_API void emu_palette_reorder(uint32_t num_indices, const uint32_t* pIndices, uint32_t num_syms, uint32_t* pRemap)
{
	const uint32_t n = num_syms;
	for (uint32_t i = 0; i < n; i++) pRemap[i] = 0;
	if (num_indices <= 1 || n == 1) return;
	std::vector<unsigned long long> keys;
	for (uint32_t i = 0; i + 1 < num_indices; i++)
	{
		const uint32_t a = pIndices[i], b = pIndices[i + 1];
		if (a != b) { keys.push_back(((unsigned long long)a << 32) | b); keys.push_back(((unsigned long long)b << 32) | a); }
	}
	std::sort(keys.begin(), keys.end());
	std::vector<unsigned long long> ukeys; std::vector<uint32_t> ucount;
	for (size_t i = 0; i < keys.size(); i++)
	{
		if (ukeys.empty() || ukeys.back() != keys[i]) { ukeys.push_back(keys[i]); ucount.push_back(1); }
		else ucount.back()++;
	}
	if (ukeys.empty()) { pRemap[0] = 1; for (uint32_t i = 2; i < n; i++) pRemap[i - 1] = i; return; }
	std::vector<uint32_t> row_ptr(n + 1, 0);
	for (size_t k = 0; k < ukeys.size(); k++) row_ptr[(uint32_t)(ukeys[k] >> 32) + 1]++;
	for (uint32_t r = 0; r < n; r++) row_ptr[r + 1] += row_ptr[r];
	size_t k0 = 0; uint32_t best_count = 0; bool have = false;
	for (size_t k = 0; k < ukeys.size(); k++)
		if ((uint32_t)(ukeys[k] >> 32) < (uint32_t)ukeys[k] && (!have || ucount[k] > best_count)) { k0 = k; best_count = ucount[k]; have = true; }
	std::vector<uint32_t> total(n, 0), n_back(n, 0), n_front(n, 0), nb_count(ukeys.size());
	std::vector<int> vpos(n, 0), nb_vpos(ukeys.size());
	std::vector<uint8_t> placed(n, 0);
	auto place = [&](uint32_t e, int vp, bool at_back) {
		for (uint32_t k = row_ptr[e]; k < row_ptr[e + 1]; k++)
		{
			const uint32_t v = (uint32_t)ukeys[k], c = ucount[k];
			if (placed[v]) continue;
			total[v] += c;
			const uint32_t slot = at_back ? (row_ptr[v] + n_back[v]++) : (row_ptr[v + 1] - 1 - n_front[v]++);
			nb_vpos[slot] = vp; nb_count[slot] = c;
		}
		placed[e] = 1; vpos[e] = vp;
	};
	place((uint32_t)(ukeys[k0] >> 32), 0, true);
	place((uint32_t)ukeys[k0], 1, true);
	int lo = 0, hi = 2;
	for (uint32_t step = 2; step < n; step++)
	{
		uint32_t e = 0; bool any = false;
		for (uint32_t u = 0; u < n; u++)
			if (!placed[u] && (!any || total[u] > total[e])) { e = u; any = true; }
		const int P = hi - lo;
		float which_side = 0.0f;
		for (uint32_t q = row_ptr[e + 1] - n_front[e]; q < row_ptr[e + 1]; q++) { const int j = nb_vpos[q] - lo, r = P + 1 - 2 * (j + 1); which_side += (float)(int)((uint32_t)r * nb_count[q]); }
		for (uint32_t q = row_ptr[e]; q < row_ptr[e] + n_back[e]; q++) { const int j = nb_vpos[q] - lo, r = P + 1 - 2 * (j + 1); which_side += (float)(int)((uint32_t)r * nb_count[q]); }
		const bool back = which_side <= 0.0f;
		place(e, back ? hi++ : --lo, back);
	}
	for (uint32_t u = 0; u < n; u++) pRemap[u] = (uint32_t)(vpos[u] - lo);
}

EMU_API void emu_etc1s_determine_selectors(const uint8_t* pBlocks, uint32_t n, const uint32_t* pColor5_inten, uint8_t* pOut, int perceptual)
{
	for (uint32_t i = 0; i < n; i++)
	{
		uint32_t px[16];
		memcpy(px, pBlocks + (size_t)i * 64, 64);
		const uint64_t v = etc1s_determine_selectors(&g_tables, perceptual != 0, px, pColor5_inten[i]);
		memcpy(pOut + (size_t)i * 8, &v, 8);
	}
}

EMU_API void emu_etc1s_refine(const uint8_t* pBlocks, uint32_t n, const b200_block_info* info, const b200_endpoint_cluster* clusters, uint32_t* out, int perceptual)
{
	for (uint32_t bi = 0; bi < n; bi++)
	{
		uint32_t px[16];
		memcpy(px, pBlocks + (size_t)bi * 64, 64);
		const b200_block_info in = info[bi];
		uint64_t best_err = UINT64_MAX;
		uint32_t best_index = 0;
		for (uint32_t k = 0; k < in.num_clusters; k++)
		{
			const b200_endpoint_cluster c = clusters[(uint32_t)in.first_cluster_ofs + k];
			if (c.etc_inten > in.cur_cluster_etc_inten) continue;
			const uint64_t e = etc1s_block_error(&g_tables, perceptual != 0, px, c.unscaled_r, c.unscaled_g, c.unscaled_b, c.etc_inten);
			if (e < best_err || (c.cluster_index == in.cur_cluster_index && e == best_err))
			{
				best_err = e; best_index = c.cluster_index;
				if (!best_err) break;
			}
		}
		out[bi] = best_index;
	}
}

// ---- UASTC RDO (bu_rdo.h): sequential host rendition of the chain the CUDA kernel runs cooperatively -------------------------
#include "../../basis_universal_b200/csrc/bu_rdo.h"
#include <map>

static block_bits load_bits(const uint8_t* p) { block_bits b; memcpy(&b.lo, p, 8); memcpy(&b.hi, p + 8, 8); return b; }
static void store_bits(uint8_t* p, const block_bits& b) { memcpy(p, &b.lo, 8); memcpy(p + 8, &b.hi, 8); }

static bool rdo_chain(uint32_t first, uint32_t last, uint8_t* blocks, const uint8_t* pixels, const rdo_params& p, std::vector<uint8_t>& modified)
{
	const bu_tables* T = &g_tables;
	const int window = (int)std::max<uint32_t>(1u, p.lz_dict_size / 16);
	std::map<std::pair<uint32_t, uint64_t>, uint32_t> history;
	for (uint32_t bi = first; bi < last; bi++)
	{
		rdo_step st;
		st.bits = load_bits(blocks + (size_t)bi * 16);
		if (!unpack_block_bits(T, st.bits, st.cur)) return false;
		if (st.cur.mode == 8) continue;
		uint32_t px[16];
		memcpy(px, pixels + (size_t)bi * 64, 64);
		st.smooth_scale = rdo_smooth_scale(p, px);
		bc7_endpoints_of(T, st.cur, st.bc7);
		const uint64_t cur_err = rdo_block_error(T, st.cur, st.bc7, px);
		st.cur_ms_err = (float)cur_err * (1.0f / 64.0f);
		st.cur_rms_err = sqrtf(st.cur_ms_err);
		mode_selector_field(st.cur.mode, st.first_sel_bit, st.total_sel_bits);
		const uint32_t n0 = std::min(64u, st.total_sel_bits);
		st.cur_sel_bits = bits_read(st.bits, st.first_sel_bit, n0);
		const auto cur_key = std::make_pair(st.first_sel_bit, st.cur_sel_bits);
		if (st.cur_rms_err >= p.skip_block_rms_thresh) { history[cur_key] = bi; continue; }

		int cur_bits;
		auto it = history.find(cur_key);
		if (it == history.end()) cur_bits = (int)((st.total_sel_bits * p.lz_literal_cost) / 100);
		else cur_bits = (int)match_cost_estimate((bi - it->second) * 16);

		float best_t = st.cur_ms_err * st.smooth_scale + (float)cur_bits * p.lambda;
		block_bits best_bits = st.bits;
		int best_index = (int)bi;
		const int first_check = std::max<int>((int)first, (int)bi - window);
		for (int pi = (int)bi - 1; pi >= first_check; --pi)
		{
			const block_bits prev = load_bits(blocks + (size_t)pi * 16);
			int match = pi;
			auto r = history.find(std::make_pair(st.first_sel_bit, bits_read(prev, st.first_sel_bit, n0)));
			if (r != history.end()) match = (int)r->second;
			float t; block_bits tb;
			if (!rdo_trial(T, p, st, px, prev, pi, match, (int)bi, t, tb)) continue;
			if (t < best_t) { best_t = t; best_index = pi; best_bits = tb; }
		}

		if (best_index != (int)bi)
		{
			candidate bc;
			if (!unpack_block_bits(T, best_bits, bc)) return false;
			if (p.endpoint_refinement && st.cur.mode == 0) rdo_refine_mode0(T, bc, px);
			best_bits = pack_without_hints(T, bc);
			store_bits(blocks + (size_t)bi * 16, best_bits);
			modified[bi] = 1;
		}
		history[std::make_pair(st.first_sel_bit, bits_read(best_bits, st.first_sel_bit, n0))] = bi;
	}
	return true;
}

EMU_API int emu_uastc_rdo(uint32_t n, uint8_t* blocks, const uint8_t* pixels, const b200_uastc_rdo_params* bp, uint32_t flags, uint32_t total_jobs)
{
	rdo_params p;
	p.lz_dict_size = bp->lz_dict_size; p.lambda = bp->lambda; p.max_allowed_rms_increase_ratio = bp->max_allowed_rms_increase_ratio;
	p.skip_block_rms_thresh = bp->skip_block_rms_thresh; p.endpoint_refinement = bp->endpoint_refinement;
	p.max_smooth_block_std_dev = bp->max_smooth_block_std_dev; p.smooth_block_max_error_scale = bp->smooth_block_max_error_scale; p.lz_literal_cost = bp->lz_literal_cost;
	std::vector<uint8_t> modified(n, 0);
	const uint32_t per_job = total_jobs ? n / total_jobs : 0;
	bool ok = true;
	if (total_jobs <= 1 || per_job <= 8) ok = rdo_chain(0, n, blocks, pixels, p, modified);
	else for (uint32_t f = 0; f < n && ok; f += per_job) ok = rdo_chain(f, std::min(n, f + per_job), blocks, pixels, p, modified);
	if (!ok) return 0;
	// deferred hint pass (uastc_recompute_hints, uastc_enc.cpp:3647)
	const int lvl = (int)(flags & 0xF);
	const level_opts o = make_level_opts(lvl);
	for (uint32_t i = 0; i < n; i++)
	{
		if (!modified[i]) continue;
		candidate c;
		if (!unpack_block_bits(&g_tables, load_bits(blocks + (size_t)i * 16), c)) return 0;
		uint32_t px[16];
		memcpy(px, pixels + (size_t)i * 64, 64);
		finish_block(&g_tables, o, lvl, flags, px, c, blocks + (size_t)i * 16);
	}
	return 1;
}

// ---- decode (b200_uastc_unpack_blocks) -----------------------------------------------------------------------------------
EMU_API int emu_uastc_unpack_blocks(const uint8_t* pUastc, uint32_t n, uint8_t* pOut64)
{
	int ok = 1;
	for (uint32_t i = 0; i < n; i++)
	{
		block_bits b;
		memcpy(&b.lo, pUastc + (size_t)i * 16, 8); memcpy(&b.hi, pUastc + (size_t)i * 16 + 8, 8);
		uint32_t px[16];
		if (!unpack_block_texels(&g_tables, b, px)) ok = 0;
		memcpy(pOut64 + (size_t)i * 64, px, 64);
	}
	return ok;
}
