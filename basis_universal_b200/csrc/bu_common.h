// bu_common.h -- host/device portability macros and tiny integer/float helpers.
//
// Every arithmetic helper here is written so that nvcc (--fmad=false, IEEE div/sqrt) and g++ (x86-64, SSE2 scalar
// math, -ffp-contract=off) produce bit-identical results: bit-exact UASTC output depends on it (hard part 1).
// The same headers are compiled twice: by nvcc into the product library, and by g++ into the host-emulation library that
// the CPU-only tests use to check the algorithm against the compiled reference without a GPU.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define BU_HD __host__ __device__
#define BU_FI __host__ __device__ __forceinline__
// Out-of-line device functions: the encoder is instruction-fetch bound when everything is inlined into one 55k-instruction
// kernel body (ncu r1: 77% of stall samples were stall_no_inst), so the big building blocks are kept as real calls.
#define BU_NI __host__ __device__ __noinline__
#else
#define BU_HD
#define BU_FI inline
#define BU_NI
#endif

// Small constant tables indexed at run time. A function-local const array would be rebuilt on the thread's stack at every call
// on the device; these live in device global memory (one copy per translation unit) and as plain statics on the host.
#if defined(__CUDACC__)
#define BU_TABLE(type, name, dims, ...) static __device__ const type name##_dev dims = __VA_ARGS__; static const type name##_host dims = __VA_ARGS__;
#else
#define BU_TABLE(type, name, dims, ...) static const type name##_host dims = __VA_ARGS__;
#endif
#if defined(__CUDA_ARCH__)
#define BU_TABLE_REF(name) name##_dev
#else
#define BU_TABLE_REF(name) name##_host
#endif

// Loops marked BU_ROLL stay rolled on the device. The encoder kernels are bound by instruction fetch (ncu: "no instruction"
// is their top stall; hot code must fit the 32 KB L1.5 instruction cache), and these loops index thread-local arrays that
// already live in local memory, so unrolling buys nothing but code size.
#if defined(__CUDA_ARCH__) && !defined(BU_NO_ROLL)
#define BU_ROLL _Pragma("unroll 1")
#else
#define BU_ROLL
#endif

namespace bu {

BU_FI int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
BU_FI int clamp255i(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
BU_FI int iabsi(int v) { return v < 0 ? -v : v; }
BU_FI int mini(int a, int b) { return a < b ? a : b; }
BU_FI int maxi(int a, int b) { return a > b ? a : b; }
BU_FI uint32_t minu(uint32_t a, uint32_t b) { return a < b ? a : b; }
BU_FI uint32_t maxu(uint32_t a, uint32_t b) { return a > b ? a : b; }
BU_FI uint64_t minu64(uint64_t a, uint64_t b) { return a < b ? a : b; }
// Same comparison forms as the reference's minimumf/maximumf/clampf/saturate (transcoder/basisu.h:136-145): NaN handling follows.
BU_FI float minf_(float a, float b) { return (a < b) ? a : b; }
BU_FI float maxf_(float a, float b) { return (a > b) ? a : b; }
BU_FI float clampf_(float v, float lo, float hi) { if (v < lo) v = lo; else if (v > hi) v = hi; return v; }
BU_FI float saturatef_(float v) { return clampf_(v, 0.0f, 1.0f); }

// float -> uint8 with x86 cvttss2si semantics for the one place the reference casts a possibly-NaN float
// (bc7enc.cpp:1660, hy == ly): NaN / out-of-int-range -> 0x80000000 -> low byte 0.
BU_FI uint8_t f2u8_x86(float v)
{
	if (!(v == v)) return 0;
	if (v >= 2147483648.0f || v < -2147483648.0f) return 0;
	return (uint8_t)(int)v;
}

// Packed RGBA8 pixel helpers (R in the low byte, as the bytes lie in memory on a little-endian machine).
BU_FI uint32_t px_c(uint32_t p, uint32_t c) { return (p >> (c * 8)) & 255u; }
BU_FI uint32_t px_make(uint32_t r, uint32_t g, uint32_t b, uint32_t a) { return r | (g << 8) | (b << 16) | (a << 24); }
BU_FI uint32_t px_set(uint32_t p, uint32_t c, uint32_t v) { return (p & ~(255u << (c * 8))) | (v << (c * 8)); }

// ASTC LDR weight interpolation, linear (non-sRGB) decode mode: transcoder_uastc.h:79 astc_interpolate(srgb=false),
// identical to bc7enc.cpp:177 astc_interpolate_linear.
BU_FI uint32_t astc_lerp(uint32_t l, uint32_t h, uint32_t w)
{
	l = (l << 8) | l;
	h = (h << 8) | h;
	return ((l * (64 - w) + h * w + 32) >> 6) >> 8;
}

// Two interpolations at once on 16-bit lanes (l, h <= 255 per lane). With t = l (64 - w) + h w <= 16320 the scalar form is
// (257 t + 32) >> 14 = (t + ((t + 32) >> 8)) >> 6 (nested floor division), and every intermediate fits its 16-bit lane.
BU_FI uint32_t astc_lerp_x2(uint32_t l2, uint32_t h2, uint32_t w)
{
	const uint32_t t = l2 * (64 - w) + h2 * w;
	return ((t + (((t + 0x00200020u) >> 8) & 0x00FF00FFu)) >> 6) & 0x00FF00FFu;
}

BU_FI uint32_t sq_diff(int a, int b) { int d = a - b; return (uint32_t)(d * d); }

// Squared distances between packed RGBA8 texels. On the device these are two instructions: per-byte |a-b| (VABSDIFF4) and
// a 4-way dot product of the differences with themselves (IDP4A); on the host the same integers are computed per channel.
#if defined(__CUDA_ARCH__)
BU_FI uint32_t dist_rgba(uint32_t p, uint32_t q) { const uint32_t d = __vabsdiffu4(p, q); return __dp4a(d, d, 0u); }
BU_FI uint32_t dist_rgb(uint32_t p, uint32_t q) { const uint32_t d = __vabsdiffu4(p, q) & 0x00FFFFFFu; return __dp4a(d, d, 0u); }
BU_FI uint32_t dist_la(uint32_t p, uint32_t q) { const uint32_t d = __vabsdiffu4(p, q) & 0xFF0000FFu; return __dp4a(d, d, 0u); }
// byte mask selects the channels that count (0x00FFFFFF rgb, 0xFFFFFFFF rgba)
BU_FI uint32_t dist_masked(uint32_t p, uint32_t q, uint32_t mask) { const uint32_t d = __vabsdiffu4(p, q) & mask; return __dp4a(d, d, 0u); }
// sum_c s16(dc[c]) * u8(p[c]) with dc packed as two s16 pairs: one DP2A each for the low and high byte pairs of p.
BU_FI int dot_s16x4_u8x4(uint32_t dc01, uint32_t dc23, uint32_t p)
{
	int r;
	asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(dc01), "r"(p), "r"(0));
	asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(dc23), "r"(p), "r"(r));
	return r;
}
BU_FI uint32_t min_u8x4(uint32_t a, uint32_t b) { return __vminu4(a, b); }
BU_FI uint32_t max_u8x4(uint32_t a, uint32_t b) { return __vmaxu4(a, b); }
#else
BU_FI uint32_t min_u8x4(uint32_t a, uint32_t b)
{
	uint32_t r = 0;
	for (uint32_t c = 0; c < 4; c++) r |= minu(px_c(a, c), px_c(b, c)) << (c * 8);
	return r;
}
BU_FI uint32_t max_u8x4(uint32_t a, uint32_t b)
{
	uint32_t r = 0;
	for (uint32_t c = 0; c < 4; c++) r |= maxu(px_c(a, c), px_c(b, c)) << (c * 8);
	return r;
}
BU_FI uint32_t dist_rgb(uint32_t p, uint32_t q)
{
	return sq_diff((int)px_c(p, 0), (int)px_c(q, 0)) + sq_diff((int)px_c(p, 1), (int)px_c(q, 1)) + sq_diff((int)px_c(p, 2), (int)px_c(q, 2));
}
BU_FI uint32_t dist_rgba(uint32_t p, uint32_t q) { return dist_rgb(p, q) + sq_diff((int)px_c(p, 3), (int)px_c(q, 3)); }
BU_FI uint32_t dist_la(uint32_t p, uint32_t q) { return sq_diff((int)px_c(p, 0), (int)px_c(q, 0)) + sq_diff((int)px_c(p, 3), (int)px_c(q, 3)); }
BU_FI uint32_t dist_masked(uint32_t p, uint32_t q, uint32_t mask) { return (mask >> 24) ? dist_rgba(p, q) : dist_rgb(p, q); }
BU_FI int dot_s16x4_u8x4(uint32_t dc01, uint32_t dc23, uint32_t p)
{
	return (int)(int16_t)(dc01 & 0xFFFF) * (int)px_c(p, 0) + (int)(int16_t)(dc01 >> 16) * (int)px_c(p, 1) +
		(int)(int16_t)(dc23 & 0xFFFF) * (int)px_c(p, 2) + (int)(int16_t)(dc23 >> 16) * (int)px_c(p, 3);
}
#endif
BU_FI uint32_t pack_s16x2(int a, int b) { return ((uint32_t)a & 0xFFFFu) | ((uint32_t)b << 16); }

} // namespace bu
