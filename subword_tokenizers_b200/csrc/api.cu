// api.cu -- library-level entry points of libswt: errors, versioning, workspace layout, pinned memory.
#include <algorithm>

#include <cstdlib>

#include "encode.cuh"

namespace swt {

static thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }

Tuning g_tune;
int g_train_timing = 0;

size_t encode_workspace_layout(uint32_t n_words, uint64_t long_bytes, void *base, EncodeWorkspace *ws) {
    Carver cv(base);
    const uint32_t n_tiles = (n_words + kTileWords - 1) / kTileWords;
    ws->n_tiles = n_tiles;
    ws->long_cursor = cv.take<unsigned long long>(2);
    // Word-type memo: rebuilt from empty by every launch; one slot per four words of the batch, at most 2^22 slots by default
    // (keys 64 MB + ids16 128 MB + ext 128 MB + tok32 128 MB; only the keys are cleared and probed by the count pass).
    // swt_tune("memo_max_log2", 10..23) moves the cap; swt_tune("memo_off", 1) disables the memo (direct-path rates).
    uint64_t slots = next_pow2(std::max<uint64_t>(n_words / 4, 1024));
    const int cap_log2 = std::min(std::max(g_tune.memo_max_log2, 10), (int)kMemoSlotBits);
    slots = std::min<uint64_t>(slots, 1ull << cap_log2);
    if (g_tune.memo_off) slots = 0;
    ws->keys = cv.take<uint4>(slots);
    ws->memo_mask = slots ? (uint32_t)(slots - 1) : 0u;
    ws->ids16 = cv.take<uint4>(2 * slots);
    ws->ext = cv.take<MemoExt>(slots);
    ws->tok32_cap = (uint32_t)std::min<uint64_t>(8 * slots, 1ull << 30);
    ws->tok32 = cv.take<uint32_t>((size_t)ws->tok32_cap + 1);
    ws->packed = cv.take<uint32_t>((size_t)n_words + 2);
    ws->tile_total = cv.take<uint32_t>((size_t)n_tiles + 1);
    ws->tile_pend = cv.take<uint2>((size_t)n_tiles + 1);
    ws->group_base = cv.take<unsigned long long>((size_t)n_tiles / 1024 + 2);
    ws->long_tiles = cv.take<uint32_t>((size_t)n_tiles / 32 + 2);
    // long words: a 16-word header + per-byte scratch each, allocated in 16-word granules.  BPE: two symbol buffers (2 words per
    // byte); WP: a 4-word segment record per boundary position of a long chunk (at most one per byte)
    ws->long_scratch_elems = long_bytes ? 4 * long_bytes + 32 * (long_bytes / 33 + 1) : 0;
    ws->long_scratch = cv.take<uint32_t>(ws->long_scratch_elems + 1);
    ws->flags = 0u;
    return cv.used();
}

}  // namespace swt

using namespace swt;

SWT_API int swt_abi_version(void) { return SWT_ABI_VERSION; }
SWT_API const char *swt_last_error(void) { return g_last_error.c_str(); }

SWT_API int swt_device_count(int *count) {
    SWT_REQUIRE(count != nullptr, "count is NULL");
    *count = 0;
    SWT_CUDA_OK(cudaGetDeviceCount(count));
    return SWT_OK;
}

SWT_API int swt_tune(const char *name, int value) {
    SWT_REQUIRE(name != nullptr, "name is NULL");
    const std::string n(name);
    if (n == "memo_max_log2") g_tune.memo_max_log2 = value;
    else if (n == "memo_off") g_tune.memo_off = value;
    else if (n == "bulk_store") g_tune.bulk_store = value;
    else if (n == "timing") g_tune.timing = value;
    else if (n == "warp_words") g_tune.warp_words = value;
    else if (n == "bpe_queue") g_tune.bpe_queue = value;
    else if (n == "split_count") g_tune.split_count = value;
    else if (n == "split_chunks") g_tune.split_chunks = value;
    else if (n == "split_warm_tiles") g_tune.split_warm_tiles = value;
    else if (n == "train_timing") g_train_timing = value;
    else { set_error("swt_tune: unknown knob " + n); return SWT_ERR_ARG; }
    return SWT_OK;
}

SWT_API size_t swt_encode_workspace_bytes(uint32_t n_words, uint64_t long_word_bytes) {
    EncodeWorkspace ws;
    return encode_workspace_layout(n_words, long_word_bytes, nullptr, &ws);
}

SWT_API int swt_host_alloc(void **ptr, size_t bytes) {
    SWT_REQUIRE(ptr != nullptr, "ptr is NULL");
    SWT_CUDA_OK(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return SWT_OK;
}
SWT_API void swt_host_free(void *ptr) { if (ptr) cudaFreeHost(ptr); }
