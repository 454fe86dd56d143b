// b200_context.cu -- context lifetime and error reporting for libbasisu_b200.so (include/basisu_b200.h).
#include "b200_internal.h"
#include <stdlib.h>

static char g_create_err[256] = "";
unsigned long long g_b200_total_launches = 0;

extern "C" int b200_device_count(void)
{
	// Load every kernel of the library when the CUDA context comes up instead of on first launch: with the default lazy loading
	// the first call of each entry point (and of each cub sort variant) pays milliseconds of module loading inside a timed stage.
	// Has to be in the environment before the runtime initialises; a value set by the user wins.
	setenv("CUDA_MODULE_LOADING", "EAGER", 0);
	int n = 0;
	const cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess) { snprintf(g_create_err, sizeof(g_create_err), "cudaGetDeviceCount: %s", cudaGetErrorString(e)); return -1; }
	return n;
}

extern "C" b200_context* b200_create_context(int device_index)
{
	const int n = b200_device_count();
	if (n <= 0) { if (n == 0) snprintf(g_create_err, sizeof(g_create_err), "no CUDA device: libbasisu_b200 has no CPU fallback"); return nullptr; }
	if (device_index < 0 || device_index >= n) { snprintf(g_create_err, sizeof(g_create_err), "device index %d out of range (0..%d)", device_index, n - 1); return nullptr; }
	cudaError_t e = cudaSetDevice(device_index);
	if (e != cudaSuccess) { snprintf(g_create_err, sizeof(g_create_err), "cudaSetDevice: %s", cudaGetErrorString(e)); return nullptr; }
	{
		// keep freed blocks in the default pool instead of returning them to the driver at every synchronisation
		cudaMemPool_t pool;
		if (cudaDeviceGetDefaultMemPool(&pool, device_index) == cudaSuccess)
		{
			unsigned long long threshold = ~0ull;
			cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
		}
	}
	b200_context* ctx = new b200_context();
	memset(ctx, 0, sizeof(*ctx));
	ctx->device = device_index;
	ctx->etc_flavour = B200_ETC1S_FLAVOUR_CPU_OPTIMIZER;
	ctx->world = 1;
	if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
		(e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess ||
		(e = cudaEventCreate(&ctx->ev_t0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev_t1)) != cudaSuccess)
	{
		snprintf(g_create_err, sizeof(g_create_err), "stream/event creation: %s", cudaGetErrorString(e));
		delete ctx;
		return nullptr;
	}
	return ctx;
}

void b200_tsvq_release(b200_context* ctx); // b200_tsvq.cu

extern "C" void b200_destroy_context(b200_context* ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	b200_tsvq_release(ctx);
	b200_comm_release(ctx);
	void* bufs[] = { ctx->d_in, ctx->d_out, ctx->d_meta, ctx->d_cands, ctx->d_errs, ctx->d_slots, ctx->d_lists, ctx->d_etc_blocks };
	for (void* p : bufs) if (p) cudaFreeAsync(p, ctx->stream);
	for (void* p : ctx->d_aux) if (p) cudaFreeAsync(p, ctx->stream);
	cudaStreamSynchronize(ctx->stream);
	cudaEventDestroy(ctx->ev0);
	cudaEventDestroy(ctx->ev1);
	cudaEventDestroy(ctx->ev_t0);
	cudaEventDestroy(ctx->ev_t1);
	for (uint32_t i = 0; i < ctx->stage_ev_count; i++) cudaEventDestroy(ctx->stage_ev[i]);
	if (ctx->copy_in) { cudaStreamDestroy(ctx->copy_in); if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out); for (int i = 0; i < 64; i++) if (ctx->pipe_ev[i]) cudaEventDestroy(ctx->pipe_ev[i]); }
	cudaStreamDestroy(ctx->stream);
	delete ctx;
}

extern "C" const char* b200_last_error(const b200_context* ctx) { return ctx ? ctx->err : g_create_err; }
extern "C" float b200_last_kernel_ms(const b200_context* ctx) { return ctx ? ctx->last_ms : 0.0f; }
extern "C" uint32_t b200_last_launch_count(const b200_context* ctx) { return ctx ? ctx->launches : 0; }

extern "C" float b200_last_stage_ms(const b200_context* ctx, uint32_t stage) { return (ctx && stage < 3) ? ctx->stage_ms[stage] : 0.0f; }

extern "C" int b200_timer_start(b200_context* ctx)
{
	if (!ctx || !ctx->activate()) return 0;
	B200_CUDA_OK(ctx, cudaEventRecord(ctx->ev_t0, ctx->stream));
	return 1;
}

extern "C" float b200_timer_stop_ms(b200_context* ctx)
{
	if (!ctx || !ctx->activate()) return -1.0f;
	float ms = -1.0f;
	if (cudaEventRecord(ctx->ev_t1, ctx->stream) != cudaSuccess) return -1.0f;
	if (cudaEventSynchronize(ctx->ev_t1) != cudaSuccess) return -1.0f;
	if (cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1) != cudaSuccess) return -1.0f;
	return ms;
}

extern "C" uint64_t b200_global_launch_count(void) { return g_b200_total_launches; }

#include <mutex>
static std::mutex g_stat_mutex;
static float g_stat_ms[B200_STAT_COUNT];
static uint32_t g_stat_launches[B200_STAT_COUNT], g_stat_calls[B200_STAT_COUNT];
void g_b200_stat_add(int stat_id, float ms, uint32_t launches)
{
	std::lock_guard<std::mutex> lk(g_stat_mutex);
	g_stat_ms[stat_id] += ms; g_stat_launches[stat_id] += launches; g_stat_calls[stat_id]++;
}
extern "C" int b200_stats_get(const b200_context* ctx, uint32_t id, float* pMs, uint32_t* pLaunches, uint32_t* pCalls)
{
	if (!ctx || id >= B200_STAT_COUNT) return 0;
	if (pMs) *pMs = ctx->stat_ms[id];
	if (pLaunches) *pLaunches = ctx->stat_launches[id];
	if (pCalls) *pCalls = ctx->stat_calls[id];
	return 1;
}
extern "C" void b200_stats_reset(b200_context* ctx)
{
	if (!ctx) return;
	for (int i = 0; i < B200_STAT_COUNT; i++) { ctx->stat_ms[i] = 0; ctx->stat_launches[i] = 0; ctx->stat_calls[i] = 0; }
}
extern "C" int b200_global_stats_get(uint32_t id, float* pMs, uint32_t* pLaunches, uint32_t* pCalls)
{
	if (id >= B200_STAT_COUNT) return 0;
	std::lock_guard<std::mutex> lk(g_stat_mutex);
	if (pMs) *pMs = g_stat_ms[id];
	if (pLaunches) *pLaunches = g_stat_launches[id];
	if (pCalls) *pCalls = g_stat_calls[id];
	return 1;
}
extern "C" void b200_global_stats_reset(void)
{
	std::lock_guard<std::mutex> lk(g_stat_mutex);
	for (int i = 0; i < B200_STAT_COUNT; i++) { g_stat_ms[i] = 0; g_stat_launches[i] = 0; g_stat_calls[i] = 0; }
}
