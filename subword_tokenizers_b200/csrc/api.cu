// api.cu -- library-level entry points of libswt: errors, versioning, workspace layout, pinned memory.
#include <algorithm>

#include <cstdlib>

#include "encode.cuh"

namespace swt {

static thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }

size_t encode_workspace_layout(uint32_t n_words, uint64_t long_bytes, void *base, EncodeWorkspace *ws) {
    Carver cv(base);
    const uint32_t n_tiles = (n_words + kTileWords - 1) / kTileWords;
    ws->n_tiles = n_tiles;
    ws->long_cursor = cv.take<unsigned long long>(1);
    // word-type memo: rebuilt from empty by every launch; sized with the batch, at most 2^20 entries (256 MB) by default.
    // Measured on the 1 GB bench stream (23 k types): 2^16 1.44+1.45 ms (count+emit), 2^18 1.22+1.38, 2^20 1.18+1.39,
    // 2^22 1.53+1.73 (the hot entries spread over 1 GB: TLB reach); a stream with 1.85 M types: 2^20 9.4 ms, 2^22 4.6 ms.
    // SWT_MEMO_MAX_LOG2 (10..23) overrides the cap for corpora with millions of word types.
    uint64_t slots = next_pow2(std::max<uint64_t>(n_words / 4, 1024));
    static const int cap_log2 = [] { const char *e = getenv("SWT_MEMO_MAX_LOG2"); const int v = e ? atoi(e) : 0; return (v >= 10 && v <= 23) ? v : 20; }();
    slots = std::min<uint64_t>(slots, 1ull << cap_log2);
    ws->memo = cv.take<MemoEntry>(slots);
    ws->memo_mask = (uint32_t)(slots - 1);
    ws->zero_bytes = cv.used();
    ws->packed = cv.take<uint32_t>((size_t)n_words + 2);
    ws->tile_total = cv.take<uint32_t>((size_t)n_tiles + 1);
    ws->group_base = cv.take<unsigned long long>((size_t)n_tiles / 1024 + 2);
    // BPE long words: 16-word header + two symbol buffers each, allocated in 16-word granules
    ws->long_scratch_elems = long_bytes ? 2 * long_bytes + 32 * (long_bytes / 33 + 1) : 0;
    ws->long_scratch = cv.take<uint32_t>(ws->long_scratch_elems + 1);
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
