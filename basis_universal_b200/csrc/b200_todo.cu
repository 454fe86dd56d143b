// b200_todo.cu -- entry points declared in include/basisu_b200.h whose kernels are not written yet. They fail loudly
// (return 0 with an error string) so a caller falls back to its own CPU code, exactly like the reference's OpenCL seam.
#include "b200_internal.h"

#define B200_TODO(name) { if (ctx) ctx->fail(name ": not implemented in this build"); return 0; }

extern "C" int b200_uastc_rdo(b200_context* ctx, uint32_t, void*, const void*, const b200_uastc_rdo_params*, uint32_t, uint32_t) B200_TODO("b200_uastc_rdo")
