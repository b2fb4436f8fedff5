// pretok.cu -- device-side pre-tokenization of FastWP.tokenize.
//
// Replaces, on raw UTF-8 text resident in HBM, what the reference does on the host before the trie walk:
//     s = text.lower() + " "                      source/wordpiece.py:248
//     ... while i < len(s) and s[i].isspace(): i += 1      :266-269 (whitespace only ever separates segments)
// i.e. Python's  text.lower().split():  the output is the packed word arena (lower-cased UTF-8 bytes of the
// whitespace-free chunks) + u32 offsets that swt_wp_encode consumes.
//
// Exactness (the tests compare with CPython on the box):
//   * whitespace = str.isspace(): the 29 code points listed in is_space_* below; the Python host verifies at start-up
//     that the running interpreter agrees with this list and refuses to use the device path otherwise;
//   * str.lower(): per-code-point table generated from the running interpreter (chr(cp).lower()), including the
//     one-to-many mapping (U+0130) and the context rule for U+03A3 (final sigma; CPython's handle_capital_sigma,
//     Objects/unicodeobject.c) with the Cased / Case_Ignorable bitmaps derived from the same interpreter.
//
// Layout: the text is analysed as 32-bit words, each with a 12-byte window (4 before, 4 after), so every decision
// (character start, whitespace, "previous character was whitespace", decode, lower) is local.  4 KiB per warp tile; the
// count pass gives a lane 16 consecutive bytes per step (128-bit loads), the write pass 4 (contiguous byte stores).  Two passes over the text (count, write) with a scan of
// the tile sums in between -- the same structure as the encode kernels, no inter-tile dependency.
#include "pretok.cuh"

namespace swt {
using namespace pt;
namespace {

template <bool kWrite, bool kBert>
__global__ void __launch_bounds__(256) pretok_kernel(PretokDev t, const uint8_t *__restrict__ text, uint64_t n, PretokWs ws,
                                                     uint8_t *__restrict__ arena, uint32_t *__restrict__ word_off, uint32_t *__restrict__ word_src, uint32_t n_words_total,
                                                     uint32_t n_bytes_total, uint32_t *status) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t tile = warp_global; tile < ws.n_tiles; tile += n_warps) {
        if constexpr (!kWrite) {
            const unsigned long long sum = count_tile<kBert>(t, text, n, tile, status);
            if (lane == 0) ws.tile_sum[tile] = sum;
        } else {
            write_tile<kBert>(t, text, n, tile, ws.group_base[tile / kGroupTiles] + ws.tile_sum[tile], arena, word_off, word_src, status);
        }
    }
    if (kWrite && blockIdx.x == 0 && threadIdx.x == 0) word_off[n_words_total] = n_bytes_total;      // closing offset
}

// in-group exclusive prefixes of the packed (words << 32 | bytes) tile sums + the group totals
__global__ void __launch_bounds__(256) pretok_scan_groups_kernel(PretokWs ws) {
    __shared__ uint32_t sh_scan[36];
    const uint32_t g = blockIdx.x, t0 = g * kGroupTiles + threadIdx.x * 4;
    unsigned long long v[4]; uint32_t sw = 0, sb = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = t0 + k < ws.n_tiles ? ws.tile_sum[t0 + k] : 0ull; sw += (uint32_t)(v[k] >> 32); sb += (uint32_t)v[k]; }
    uint32_t tw, tb;
    uint32_t ew = block_exclusive_scan(sw, sh_scan, &tw);
    uint32_t eb = block_exclusive_scan(sb, sh_scan, &tb);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (t0 + k < ws.n_tiles) ws.tile_sum[t0 + k] = ((unsigned long long)ew << 32) | eb;
        ew += (uint32_t)(v[k] >> 32); eb += (uint32_t)v[k];
    }
    if (threadIdx.x == 0) ws.group_base[g] = ((unsigned long long)tw << 32) | tb;
}
// exclusive scan of the group totals (one thread: at most a few thousand groups) + the grand totals
__global__ void pretok_scan_top_kernel(PretokWs ws, uint32_t *status) {
    unsigned long long words = 0, bytes = 0;
    for (uint32_t g = 0; g < ws.n_groups; ++g) {
        const unsigned long long v = ws.group_base[g];
        ws.group_base[g] = (words << 32) | (bytes & 0xFFFFFFFFull);
        words += v >> 32; bytes += v & 0xFFFFFFFFull;
    }
    status[kPtWords] = (uint32_t)words; status[kPtBytesLo] = (uint32_t)bytes; status[kPtBytesHi] = (uint32_t)(bytes >> 32);
    if (bytes >= 0xFFFFFFC0ull || words >= 0xFFFFFFFEull) status[kPtCode] = SWT_ERR_CAPACITY;
}

int pretok_grid(uint32_t n_tiles) { return (int)std::min<uint32_t>((n_tiles + 7) / 8, kNumSMs * 8); }

}  // namespace
}  // namespace swt

using namespace swt;

SWT_API int swt_pretok_create(const uint32_t *lower_map, uint32_t n_lower, const uint32_t *multi, uint32_t n_multi,
                              const uint8_t *cased_bitmap, const uint8_t *ignorable_bitmap, int mode, int device, swt_pretok **out) {
    SWT_REQUIRE(mode == SWT_PRETOK_PYTHON_SPLIT || mode == SWT_PRETOK_BERT, "unknown pre-tokenizer mode");
    SWT_REQUIRE(out != nullptr && lower_map != nullptr && n_lower > 0, "NULL argument");
    SWT_REQUIRE((cased_bitmap == nullptr) == (ignorable_bitmap == nullptr), "cased / ignorable bitmaps come together");
    SWT_REQUIRE(n_multi == 0 || multi != nullptr, "multi is NULL");
    SWT_CUDA_OK(cudaSetDevice(device));
    swt_pretok *p = new swt_pretok();
    p->device = device; p->mode = mode; p->n_lower = n_lower; p->n_multi = n_multi;
    p->d_lower = nullptr; p->d_multi = nullptr; p->d_cased = p->d_ignorable = nullptr;
    const size_t bm = 0x110000 / 8;
    cudaError_t e = cudaMalloc(&p->d_lower, (size_t)n_lower * 4);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_lower, lower_map, (size_t)n_lower * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_multi, ((size_t)n_multi + 4) * 4);
    if (e == cudaSuccess) e = cudaMemset(p->d_multi, 0, ((size_t)n_multi + 4) * 4);
    if (e == cudaSuccess && n_multi) e = cudaMemcpy(p->d_multi, multi, (size_t)n_multi * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && cased_bitmap) {
        e = cudaMalloc(&p->d_cased, bm);
        if (e == cudaSuccess) e = cudaMalloc(&p->d_ignorable, bm);
        if (e == cudaSuccess) e = cudaMemcpy(p->d_cased, cased_bitmap, bm, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(p->d_ignorable, ignorable_bitmap, bm, cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        set_error(std::string("swt_pretok_create: ") + cudaGetErrorString(e));
        swt_pretok_destroy(p);
        return SWT_ERR_CUDA;
    }
    *out = p;
    return SWT_OK;
}

namespace swt {
int pretok_mode(const swt_pretok *p) { return p->mode; }
uint32_t pretok_small_max_bytes() { return kSmallTiles * kTileBytes; }
PretokDev pretok_dev_view(const swt_pretok *p) { return dev_view(p); }
}  // namespace swt

SWT_API void swt_pretok_destroy(swt_pretok *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    cudaFree(p->d_lower); cudaFree(p->d_multi); cudaFree(p->d_cased); cudaFree(p->d_ignorable);
    delete p;
}

SWT_API size_t swt_pretok_workspace_bytes(uint64_t n_text_bytes) {
    PretokWs ws;
    return pretok_layout(n_text_bytes, nullptr, &ws);
}

SWT_API int swt_pretok_count(const swt_pretok *p, const uint8_t *d_text, uint64_t n_bytes, void *d_workspace, size_t workspace_bytes,
                             uint32_t *d_status, void *stream) {
    SWT_REQUIRE(p && d_status && d_workspace, "NULL argument");
    SWT_REQUIRE(n_bytes == 0 || d_text, "d_text is NULL");
    SWT_REQUIRE(((uintptr_t)d_text & 15) == 0, "d_text must be 16-byte aligned");
    SWT_REQUIRE(n_bytes < 0xFFFFFF00ull, "text must be < 4 GiB per call");
    cudaStream_t st = (cudaStream_t)stream;
    PretokWs ws;
    if (pretok_layout(n_bytes, d_workspace, &ws) > workspace_bytes) { set_error("pretok workspace too small"); return SWT_ERR_CAPACITY; }
    SWT_CUDA_OK(cudaMemsetAsync(d_status, 0, 8 * sizeof(uint32_t), st));
    if (n_bytes == 0) return SWT_OK;
    if (p->mode == SWT_PRETOK_BERT) pretok_kernel<false, true><<<pretok_grid(ws.n_tiles), 256, 0, st>>>(dev_view(p), d_text, n_bytes, ws, nullptr, nullptr, nullptr, 0, 0, d_status);
    else pretok_kernel<false, false><<<pretok_grid(ws.n_tiles), 256, 0, st>>>(dev_view(p), d_text, n_bytes, ws, nullptr, nullptr, nullptr, 0, 0, d_status);
    pretok_scan_groups_kernel<<<ws.n_groups, 256, 0, st>>>(ws);
    pretok_scan_top_kernel<<<1, 1, 0, st>>>(ws, d_status);
    SWT_CUDA_OK(cudaGetLastError());
    return SWT_OK;
}

SWT_API int swt_pretok_write(const swt_pretok *p, const uint8_t *d_text, uint64_t n_bytes, void *d_workspace, size_t workspace_bytes,
                             uint8_t *d_arena_out, uint64_t arena_cap, uint32_t *d_word_off_out, uint32_t *d_word_src_out, uint64_t word_cap,
                             uint32_t n_words, uint64_t n_out_bytes, uint32_t *d_status, void *stream) {
    SWT_REQUIRE(p && d_status && d_workspace && d_word_off_out, "NULL argument");
    SWT_REQUIRE(n_out_bytes == 0 || d_arena_out, "d_arena_out is NULL");
    SWT_REQUIRE(((uintptr_t)d_text & 15) == 0, "d_text must be 16-byte aligned");
    SWT_REQUIRE(arena_cap >= n_out_bytes && word_cap >= (uint64_t)n_words + 1, "output capacity below the counts of swt_pretok_count");
    cudaStream_t st = (cudaStream_t)stream;
    PretokWs ws;
    if (pretok_layout(n_bytes, d_workspace, &ws) > workspace_bytes) { set_error("pretok workspace too small"); return SWT_ERR_CAPACITY; }
    if (n_bytes == 0) { SWT_CUDA_OK(cudaMemsetAsync(d_word_off_out, 0, sizeof(uint32_t), st)); return SWT_OK; }
    if (p->mode == SWT_PRETOK_BERT)
        pretok_kernel<true, true><<<pretok_grid(ws.n_tiles), 256, 0, st>>>(dev_view(p), d_text, n_bytes, ws, d_arena_out, d_word_off_out, d_word_src_out, n_words, (uint32_t)n_out_bytes, d_status);
    else
        pretok_kernel<true, false><<<pretok_grid(ws.n_tiles), 256, 0, st>>>(dev_view(p), d_text, n_bytes, ws, d_arena_out, d_word_off_out, d_word_src_out, n_words, (uint32_t)n_out_bytes, d_status);
    SWT_CUDA_OK(cudaGetLastError());
    return SWT_OK;
}
