"""ctypes binding of libbasisu_b200.so -- the same C ABI (include/basisu_b200.h) a C++ host links against.
There is no CPU fallback: a missing library or missing GPU raises."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# BASISU_B200_LIB selects another build of the same library (tools/build_variant.sh); there is still no CPU fallback.
LIB_PATH = os.environ.get("BASISU_B200_LIB") or os.path.join(HERE, "libbasisu_b200.so")

_lib = None


class B200Error(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(f"{LIB_PATH} is missing: build it with `python -m basis_universal_b200.build` (there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        vp, u32, i32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int
        L.b200_device_count.restype = i32
        L.b200_create_context.restype = vp
        L.b200_create_context.argtypes = [i32]
        L.b200_destroy_context.argtypes = [vp]
        L.b200_last_error.restype = ctypes.c_char_p
        L.b200_last_error.argtypes = [vp]
        L.b200_last_kernel_ms.restype = ctypes.c_float
        L.b200_last_kernel_ms.argtypes = [vp]
        L.b200_last_launch_count.restype = u32
        L.b200_last_launch_count.argtypes = [vp]
        L.b200_global_launch_count.restype = ctypes.c_uint64
        L.b200_last_stage_ms.restype = ctypes.c_float
        L.b200_last_stage_ms.argtypes = [vp, u32]
        L.b200_timer_start.restype = i32
        L.b200_timer_start.argtypes = [vp]
        L.b200_timer_stop_ms.restype = ctypes.c_float
        L.b200_timer_stop_ms.argtypes = [vp]
        for name in ("b200_uastc_encode_blocks", "b200_uastc_encode_blocks_device"):
            f = getattr(L, name)
            f.restype = i32
            f.argtypes = [vp, vp, u32, vp, u32]
        L.b200_uastc_rdo.restype = i32
        L.b200_uastc_rdo.argtypes = [vp, u32, vp, vp, vp, u32, u32]
        for name in ("b200_uastc_rdo_batch", "b200_uastc_rdo_batch_device"):
            f = getattr(L, name)
            f.restype = i32
            f.argtypes = [vp, u32, vp, vp, vp, vp, u32, u32]
        L.b200_uastc_encode_rdo_blocks.restype = i32
        L.b200_uastc_encode_rdo_blocks.argtypes = [vp, u32, vp, vp, vp, u32, vp, u32]
        L.b200_etc1s_endpoint_histogram.restype = i32
        L.b200_etc1s_endpoint_histogram.argtypes = [vp, vp, u32, vp]
        L.b200_etc1s_endpoint_histogram_device.restype = i32
        L.b200_etc1s_endpoint_histogram_device.argtypes = [vp, vp, u32, vp]
        L.b200_etc1s_selector_training.restype = i32
        L.b200_etc1s_selector_training.argtypes = [vp, vp, u32, i32, vp, vp]
        L.b200_etc1s_selector_training_device.restype = i32
        L.b200_etc1s_selector_training_device.argtypes = [vp, vp, u32, i32, vp, vp]
        sz = ctypes.c_size_t
        L.b200_uastc_encode_image.restype = i32
        L.b200_uastc_encode_image.argtypes = [vp, vp, u32, u32, sz, vp, u32]
        L.b200_extract_source_blocks.restype = i32
        L.b200_extract_source_blocks.argtypes = [vp, vp, u32, u32, sz, vp]
        L.b200_extract_source_blocks_device.restype = i32
        L.b200_extract_source_blocks_device.argtypes = [vp, vp, u32, u32, sz, vp]
        L.b200_uastc_unpack_blocks.restype = i32
        L.b200_uastc_unpack_blocks.argtypes = [vp, vp, u32, vp]
        L.b200_uastc_unpack_blocks_device.restype = i32
        L.b200_uastc_unpack_blocks_device.argtypes = [vp, vp, u32, vp]
        L.b200_etc1_unpack_blocks.restype = i32
        L.b200_etc1_unpack_blocks.argtypes = [vp, vp, u32, vp]
        L.b200_etc1_unpack_blocks_device.restype = i32
        L.b200_etc1_unpack_blocks_device.argtypes = [vp, vp, u32, vp]
        L.b200_block_metrics_device.restype = i32
        L.b200_block_metrics_device.argtypes = [vp, vp, vp, u32, u32, vp]
        L.b200_etc1s_set_flavour.restype = i32
        L.b200_etc1s_set_flavour.argtypes = [vp, i32]
        L.b200_etc1s_set_pixel_blocks.restype = i32
        L.b200_etc1s_set_pixel_blocks.argtypes = [vp, u32, vp]
        L.b200_etc1s_encode_blocks.restype = i32
        L.b200_etc1s_encode_blocks.argtypes = [vp, vp, i32, u32]
        L.b200_etc1s_encode_pixel_clusters.restype = i32
        L.b200_etc1s_encode_pixel_clusters.argtypes = [vp, vp, u32, vp, ctypes.c_uint64, vp, vp, i32, u32]
        L.b200_etc1s_refine_endpoint_clusterization.restype = i32
        L.b200_etc1s_refine_endpoint_clusterization.argtypes = [vp, vp, u32, vp, vp, vp, i32]
        L.b200_etc1s_find_optimal_selector_clusters_for_each_block.restype = i32
        L.b200_etc1s_find_optimal_selector_clusters_for_each_block.argtypes = [vp, vp, u32, vp, vp, vp, i32]
        L.b200_etc1s_determine_selectors.restype = i32
        L.b200_etc1s_determine_selectors.argtypes = [vp, vp, vp, i32]
        L.b200_comm_unique_id.restype = i32
        L.b200_comm_unique_id.argtypes = [vp]
        L.b200_comm_init.restype = i32
        L.b200_comm_init.argtypes = [vp, i32, i32, vp]
        L.b200_shard_range.restype = None
        L.b200_shard_range.argtypes = [u32, u32, u32, vp, vp]
        L.b200_comm_rank.restype = i32
        L.b200_comm_rank.argtypes = [vp]
        L.b200_comm_world.restype = i32
        L.b200_comm_world.argtypes = [vp]
        L.b200_comm_allreduce_u32_device.restype = i32
        L.b200_comm_allreduce_u32_device.argtypes = [vp, vp, sz]
        L.b200_comm_stats.restype = i32
        L.b200_comm_stats.argtypes = [vp, vp, vp, vp]
        L.b200_comm_last_error.restype = ctypes.c_char_p
        L.b200_stats_get.restype = i32
        L.b200_stats_get.argtypes = [vp, u32, vp, vp, vp]
        L.b200_stats_reset.argtypes = [vp]
        L.b200_global_stats_get.restype = i32
        L.b200_global_stats_get.argtypes = [u32, vp, vp, vp]
        L.b200_tsvq_generate.restype = i32
        L.b200_tsvq_generate.argtypes = [vp, u32, u32, vp, sz, sz, u32, u32, u32, i32, vp]
        L.b200_etc1s_encode_endpoint_clusters.restype = i32
        L.b200_etc1s_encode_endpoint_clusters.argtypes = [vp, vp, u32, vp, vp, i32, u32]
        L.b200_etc1s_optimize_selector_codebook.restype = i32
        L.b200_etc1s_optimize_selector_codebook.argtypes = [vp, vp, u32, vp, vp, vp, i32]
        L.b200_etc1s_reoptimize_endpoint_clusters.restype = i32
        L.b200_etc1s_reoptimize_endpoint_clusters.argtypes = [vp, u32, vp, vp, vp, vp, vp, vp, vp, i32, u32]
        L.b200_etc1s_subblock_errors.restype = i32
        L.b200_etc1s_subblock_errors.argtypes = [vp, vp, vp, i32]
        L.b200_etc1s_backend_endpoint_prediction.restype = i32
        L.b200_etc1s_backend_endpoint_prediction.argtypes = [vp, u32, vp, vp, u32, vp, ctypes.c_float, i32, vp, vp]
        L.b200_image_resample_rgba8.restype = i32
        L.b200_image_resample_rgba8.argtypes = [vp, vp, u32, u32, sz, vp, u32, u32, sz, vp, vp, vp, vp, u32, u32, vp, vp]
        L.b200_palette_reorder.restype = i32
        L.b200_palette_reorder.argtypes = [vp, u32, vp, u32, vp]
        _lib = L
    return _lib


STAT_NAMES = ["etc1s_encode_blocks", "etc1s_endpoint_clusters", "etc1s_refine", "etc1s_determine_selectors", "etc1s_find_selector_clusters",
              "etc1s_selector_codebook", "tsvq", "uastc_encode", "uastc_rdo", "etc1s_reoptimize_clusters", "etc1s_backend_prediction"]


def global_stats():
    """{family: (kernel_ms, launches, calls)} over every context of this process (b200_global_stats_get)."""
    L = lib()
    out = {}
    for i, name in enumerate(STAT_NAMES):
        ms, la, ca = ctypes.c_float(0), ctypes.c_uint32(0), ctypes.c_uint32(0)
        L.b200_global_stats_get(i, ctypes.byref(ms), ctypes.byref(la), ctypes.byref(ca))
        out[name] = (float(ms.value), int(la.value), int(ca.value))
    return out


EXPORTS = [
    "b200_device_count", "b200_create_context", "b200_destroy_context", "b200_last_error",
    "b200_uastc_encode_blocks", "b200_uastc_encode_blocks_device", "b200_uastc_encode_image", "b200_uastc_rdo", "b200_uastc_rdo_batch", "b200_uastc_rdo_batch_device", "b200_uastc_encode_rdo_blocks",
    "b200_etc1s_set_flavour", "b200_etc1s_set_pixel_blocks", "b200_etc1s_endpoint_histogram", "b200_etc1s_endpoint_histogram_device",
    "b200_etc1s_selector_training", "b200_etc1s_selector_training_device",
    "b200_extract_source_blocks", "b200_extract_source_blocks_device", "b200_uastc_unpack_blocks", "b200_uastc_unpack_blocks_device", "b200_block_metrics_device", "b200_etc1_unpack_blocks", "b200_etc1_unpack_blocks_device", "b200_etc1s_encode_blocks", "b200_etc1s_encode_pixel_clusters",
    "b200_etc1s_refine_endpoint_clusterization", "b200_etc1s_find_optimal_selector_clusters_for_each_block",
    "b200_etc1s_determine_selectors", "b200_last_kernel_ms", "b200_last_launch_count", "b200_last_stage_ms",
    "b200_timer_start", "b200_timer_stop_ms", "b200_global_launch_count",
    "b200_comm_unique_id", "b200_comm_init", "b200_shard_range", "b200_comm_rank", "b200_comm_world", "b200_comm_allreduce_u32_device", "b200_comm_stats", "b200_comm_last_error",
    "b200_stats_get", "b200_stats_reset", "b200_global_stats_get", "b200_global_stats_reset", "b200_tsvq_generate", "b200_etc1s_encode_endpoint_clusters", "b200_etc1s_optimize_selector_codebook",
    "b200_etc1s_reoptimize_endpoint_clusters", "b200_etc1s_subblock_errors", "b200_etc1s_backend_endpoint_prediction", "b200_image_resample_rgba8", "b200_palette_reorder",
]
