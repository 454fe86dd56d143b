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
