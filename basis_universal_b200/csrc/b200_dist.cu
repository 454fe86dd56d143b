// b200_dist.cu -- multi-GPU plumbing of libbasisu_b200.so: one process per GPU, NCCL over NVLink/NVSwitch.
//
// The ETC1S frontend shards naturally: every per-block stage splits by block-row ranges, every per-cluster stage
// by clusters. The host logic between the stages (the reference's frontend, unchanged) runs replicated on every rank on identical
// data, so the only exchange is the MERGE OF STAGE OUTPUTS: each rank computes its share of a stage's output array into a
// zeroed device buffer and one ncclAllReduce(sum, u32) over the array gives every rank the complete result (each element is
// written by exactly one rank, so the sum is a merge; 4-8 bytes per block or cluster, <= 32 MB for an 8192^2 image).
// UASTC needs no collective at all.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy a torch process has already loaded, else the system one), so a
// single-GPU host has no NCCL dependency.
#include "b200_internal.h"
#include <dlfcn.h>
#include <nccl.h> // types and enums only; every function is looked up with dlsym

namespace
{
	struct nccl_api
	{
		void* lib = nullptr;
		ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
		ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
		ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
		ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
		const char* (*GetErrorString)(ncclResult_t) = nullptr;
		bool load(char* err, size_t err_size)
		{
			if (lib) return true;
			const char* override_path = getenv("B200_NCCL_LIB");
			lib = dlopen(override_path ? override_path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
			if (!lib) { snprintf(err, err_size, "dlopen(libnccl.so.2): %s", dlerror()); return false; }
			GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
			CommInitRank = reinterpret_cast<decltype(CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
			CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
			AllReduce = reinterpret_cast<decltype(AllReduce)>(dlsym(lib, "ncclAllReduce"));
			GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
			if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !GetErrorString) { snprintf(err, err_size, "libnccl.so.2 lacks a required symbol"); lib = nullptr; return false; }
			return true;
		}
	};
	nccl_api g_nccl;
	char g_dist_err[256] = "";

	// An NCCL unique id makes ONE communicator. A compressor creates a context per image (basis_compressor does), so the
	// communicator is kept for the life of the process and every context initialised with the same id attaches to it.
	struct comm_cache { bool valid = false; uint8_t id[128]; ncclComm_t comm = nullptr; int rank = 0, world = 1, device = -1; };
	comm_cache g_comm;
}

static_assert(sizeof(ncclUniqueId) == 128, "b200_comm_unique_id hands out 128 bytes");

extern "C" int b200_comm_unique_id(uint8_t* pId128)
{
	if (!pId128) return 0;
	if (!g_nccl.load(g_dist_err, sizeof(g_dist_err))) return 0;
	ncclUniqueId id;
	const ncclResult_t r = g_nccl.GetUniqueId(&id);
	if (r != ncclSuccess) { snprintf(g_dist_err, sizeof(g_dist_err), "ncclGetUniqueId: %s", g_nccl.GetErrorString(r)); return 0; }
	memcpy(pId128, &id, 128);
	return 1;
}

extern "C" int b200_comm_init(b200_context* ctx, int rank, int world, const uint8_t* pId128)
{
	if (!ctx) return 0;
	if (!ctx->activate()) return 0;
	if (world < 1 || rank < 0 || rank >= world || !pId128) { ctx->fail("b200_comm_init: bad rank/world/id"); return 0; }
	if (ctx->comm) { ctx->fail("b200_comm_init: the context already has a communicator"); return 0; }
	if (world == 1) { ctx->rank = 0; ctx->world = 1; return 1; }
	if (!g_nccl.load(ctx->err, sizeof(ctx->err))) return 0;
	if (g_comm.valid && !memcmp(g_comm.id, pId128, 128))
	{
		if (g_comm.rank != rank || g_comm.world != world || g_comm.device != ctx->device) { ctx->fail("b200_comm_init: this id already made a communicator with another rank/world/device"); return 0; }
		ctx->comm = g_comm.comm; ctx->rank = rank; ctx->world = world;
		return 1;
	}
	ncclUniqueId id;
	memcpy(&id, pId128, 128);
	ncclComm_t comm = nullptr;
	const ncclResult_t r = g_nccl.CommInitRank(&comm, world, id, rank);
	if (r != ncclSuccess) { snprintf(ctx->err, sizeof(ctx->err), "ncclCommInitRank: %s", g_nccl.GetErrorString(r)); return 0; }
	if (g_comm.valid) g_nccl.CommDestroy(g_comm.comm); // a new id replaces the cached communicator (no context uses the old one any more in a well-formed host)
	g_comm.valid = true; memcpy(g_comm.id, pId128, 128); g_comm.comm = comm; g_comm.rank = rank; g_comm.world = world; g_comm.device = ctx->device;
	ctx->comm = comm; ctx->rank = rank; ctx->world = world;
	return 1;
}

void b200_comm_release(b200_context* ctx)
{
	if (ctx && ctx->comm) { ctx->comm = nullptr; ctx->world = 1; ctx->rank = 0; } // the communicator itself lives in g_comm until the process ends or a new id arrives
}

// [first, last) of `rank`'s share of n per-block units: equal contiguous ranges of ceil(n / world), the tail ranks short or empty.
extern "C" void b200_shard_range(uint32_t n, uint32_t rank, uint32_t world, uint32_t* pFirst, uint32_t* pLast)
{
	uint32_t first = 0, last = n;
	if (world > 1)
	{
		const uint32_t per = (n + world - 1) / world;
		const uint64_t f = (uint64_t)rank * per;
		first = (f > n) ? n : (uint32_t)f;
		last = (n - first > per) ? first + per : n;
	}
	if (pFirst) *pFirst = first;
	if (pLast) *pLast = last;
}

extern "C" int b200_comm_rank(const b200_context* ctx) { return ctx ? ctx->rank : 0; }
extern "C" int b200_comm_world(const b200_context* ctx) { return (ctx && ctx->world > 1) ? ctx->world : 1; }

// In-place SUM all-reduce of `count` u32 on the context's stream (enqueued; the caller synchronises). No-op for one rank.
bool b200_merge_u32(b200_context* ctx, void* d_buf, size_t count)
{
	if (!ctx->comm || ctx->world <= 1 || !count) return true;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	const bool timed = cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess;
	if (timed) cudaEventRecord(e0, ctx->stream);
	const ncclResult_t r = g_nccl.AllReduce(d_buf, d_buf, count, ncclUint32, ncclSum, static_cast<ncclComm_t>(ctx->comm), ctx->stream);
	if (r != ncclSuccess) { snprintf(ctx->err, sizeof(ctx->err), "ncclAllReduce: %s", g_nccl.GetErrorString(r)); return false; }
	if (timed)
	{
		cudaEventRecord(e1, ctx->stream);
		if (cudaEventSynchronize(e1) == cudaSuccess) { float ms = 0; if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) { ctx->comm_ms += ms; ctx->comm_bytes += count * 4; ctx->comm_calls++; } }
	}
	if (e0) cudaEventDestroy(e0);
	if (e1) cudaEventDestroy(e1);
	return true;
}

extern "C" int b200_comm_allreduce_u32_device(b200_context* ctx, void* dBuf, size_t count)
{
	if (!ctx) return 0;
	if (!ctx->activate()) return 0;
	if (!b200_merge_u32(ctx, dBuf, count)) return 0;
	B200_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
	return 1;
}

extern "C" int b200_comm_stats(const b200_context* ctx, float* pMs, uint64_t* pBytes, uint32_t* pCalls)
{
	if (!ctx) return 0;
	if (pMs) *pMs = ctx->comm_ms;
	if (pBytes) *pBytes = ctx->comm_bytes;
	if (pCalls) *pCalls = ctx->comm_calls;
	return 1;
}

extern "C" const char* b200_comm_last_error(void) { return g_dist_err; }
