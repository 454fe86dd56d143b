/* oracle/etc1s_port.c -- TEST INFRASTRUCTURE: plain-C restatement ("port") of the integer ETC1S per-block routines on the
 * hot path.  It exists so the parity tests have a third, independently written opinion next to the compiled reference
 * (oracle/_ref/libbasisu_ref.so) and the reference's OpenCL kernels compiled for the host (oracle/_ref/libocl_ref.so);
 * it is pinned against both by tests/test_etc1s_cpu.py::test_plain_c_port_*.  Never linked into the product.
 *
 *   port_color_distance          basisu::color_distance            encoder/basisu_enc.h:1141-1195
 *   port_block_colors5           etc_block::get_block_colors5      encoder/basisu_etc.h:672-690
 *   port_determine_selectors     etc_block::determine_selectors    encoder/basisu_etc.h:374-430 (ETC1S: one colour, one table)
 *   port_block_error             inner loop of basisu_frontend::refine_endpoint_clusterization, encoder/basisu_frontend.cpp:1846-1890
 */
#include <stdint.h>
#include <string.h>

static const int k_inten[8][4] = { /* g_etc1_inten_tables, encoder/basisu_etc.cpp:304 */
	{ -8, -2, 2, 8 }, { -17, -5, 5, 17 }, { -29, -9, 9, 29 }, { -42, -13, 13, 42 },
	{ -60, -18, 18, 60 }, { -80, -24, 24, 80 }, { -106, -33, 33, 106 }, { -183, -47, 47, 183 } };
static const uint8_t k_selector_index_to_etc1[4] = { 3, 2, 0, 1 }; /* etc.cpp:311 */

static int clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

uint32_t port_color_distance(int perceptual, const uint8_t* a, const uint8_t* b)
{
	int dr = a[0] - b[0], dg = a[1] - b[1], db = a[2] - b[2];
	if (perceptual)
	{
		int dl = dr * 14 + dg * 45 + db * 5;
		int dcr = dr * 64 - dl, dcb = db * 64 - dl;
		return ((uint32_t)(dl * dl) >> 5) + ((((uint32_t)(dcr * dcr) >> 5) * 26u) >> 7) + ((((uint32_t)(dcb * dcb) >> 5) * 3u) >> 7);
	}
	return (uint32_t)(dr * dr + dg * dg + db * db);
}

void port_block_colors5(const uint8_t* rgb5, uint32_t inten, uint8_t colors[4][4])
{
	int c, s;
	for (s = 0; s < 4; s++)
	{
		for (c = 0; c < 3; c++)
		{
			int v = (rgb5[c] << 3) | (rgb5[c] >> 2);
			colors[s][c] = (uint8_t)clamp255(v + k_inten[inten][s]);
		}
		colors[s][3] = 255;
	}
}

/* One ETC1S block (flip = 1, diff = 1, delta 0) with the lowest-error selector per texel, ties to the lowest selector. */
void port_determine_selectors(const uint8_t* pixels64, const uint8_t* rgb5_inten, int perceptual, uint8_t out[8])
{
	uint8_t colors[4][4];
	uint32_t lsb = 0, msb = 0, i, s;
	port_block_colors5(rgb5_inten, rgb5_inten[3], colors);
	for (i = 0; i < 16; i++)
	{
		uint32_t best = port_color_distance(perceptual, pixels64 + i * 4, colors[0]), bs = 0;
		for (s = 1; s < 4; s++)
		{
			uint32_t e = port_color_distance(perceptual, pixels64 + i * 4, colors[s]);
			if (e < best) { best = e; bs = s; }
		}
		{
			uint32_t raw = k_selector_index_to_etc1[bs], bit = (i & 3) * 4 + (i >> 2); /* etc.h:264: bit index = x*4+y */
			lsb |= (raw & 1) << bit; msb |= (raw >> 1) << bit;
		}
	}
	out[0] = (uint8_t)(rgb5_inten[0] << 3); out[1] = (uint8_t)(rgb5_inten[1] << 3); out[2] = (uint8_t)(rgb5_inten[2] << 3);
	out[3] = (uint8_t)((rgb5_inten[3] << 5) | (rgb5_inten[3] << 2) | 3);
	out[4] = (uint8_t)(msb >> 8); out[5] = (uint8_t)msb; out[6] = (uint8_t)(lsb >> 8); out[7] = (uint8_t)lsb;
}

uint64_t port_block_error(const uint8_t* pixels64, const uint8_t* rgb5, uint32_t inten, int perceptual)
{
	uint8_t colors[4][4];
	uint64_t total = 0;
	uint32_t i, s;
	port_block_colors5(rgb5, inten, colors);
	for (i = 0; i < 16; i++)
	{
		uint32_t best = port_color_distance(perceptual, pixels64 + i * 4, colors[0]);
		for (s = 1; s < 4; s++)
		{
			uint32_t e = port_color_distance(perceptual, pixels64 + i * 4, colors[s]);
			if (e < best) best = e;
		}
		total += best;
	}
	return total;
}
