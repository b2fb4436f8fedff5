// pipeline.cu -- host-buffer entry point: stages a host-resident corpus through the GPU in batches.
//
// Five slots, one CUDA stream each; per batch k on slot k%5:  H2D(arena slice, offset slice) -> encode
// kernel -> D2H(token ids, token offsets).  H2D of batch k+1 and D2H of batch k-1 overlap the kernel of
// batch k (PCIe is full duplex).  The kernel of batch k is launched once the token total of batch k-1 is
// known, so that token offsets come out global and ids land compactly in the caller's buffer.
#include <algorithm>
#include <cstring>
#include <vector>

#include "encode.cuh"

struct swt_bpe_table;
struct swt_wp_trie;
namespace swt {
int pretok_mode(const swt_pretok *p);
uint32_t pretok_small_max_bytes();
pt::PretokDev pretok_dev_view(const swt_pretok *p);
int bpe_small_launch(const swt_bpe_table *t, int naive, const pt::PretokDev &pd, bool bert, const uint8_t *h_text, const uint8_t *d_text, uint32_t n,
                     const SmallArgs &a, cudaStream_t st);
int wp_small_launch(const swt_wp_trie *t, int naive, const pt::PretokDev &pd, bool bert, const uint8_t *h_text, const uint8_t *d_text, uint32_t n,
                     const SmallArgs &a, cudaStream_t st);
int bpe_encode_launch(const swt_bpe_table *t, const uint8_t *d_arena, const uint32_t *d_word_off, uint32_t n_words,
                      uint64_t long_word_bytes, uint32_t *d_out_ids, uint64_t out_cap, uint32_t *d_out_tok_off, uint32_t tok_base,
                      void *d_workspace, size_t workspace_bytes, uint32_t *d_status, cudaStream_t st);
int wp_encode_launch(const swt_wp_trie *t, const uint8_t *d_arena, const uint32_t *d_word_off, uint32_t n_words, uint64_t long_word_bytes,
                     uint32_t *d_out_ids, uint64_t out_cap, uint32_t *d_out_tok_off, uint32_t tok_base,
                     void *d_workspace, size_t workspace_bytes, uint32_t *d_status, cudaStream_t st);
}

// inside the batch loops: a failed CUDA call leaves the loop (rc set) so that the slots in flight are drained before returning
#define SWT_CUDA_BRK(expr)                                                                        \
    {                                                                                             \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            ::swt::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                 \
            rc = SWT_ERR_CUDA;                                                                    \
            break;                                                                                \
        }                                                                                         \
    }

namespace {
constexpr int kSlots = 5;
constexpr uint64_t kLowerGrowthNum = 3, kLowerGrowthDen = 2;   // str.lower() grows UTF-8 text by at most 3/2 (2-byte -> 3-byte)
struct Slot {
    uint8_t *d_arena = nullptr; uint32_t *d_off = nullptr, *d_ids = nullptr, *d_tok_off = nullptr, *d_status = nullptr;
    uint16_t *d_ids16 = nullptr;              // 16-bit output mode
    uint8_t *d_text = nullptr; void *d_ptws = nullptr; size_t ptws_bytes = 0;   // raw-text mode
    void *d_ws = nullptr; size_t ws_bytes = 0;
    uint32_t *h_status = nullptr;             // pinned
    cudaStream_t stream = nullptr;
    cudaEvent_t kernel_done = nullptr, d2h_done = nullptr;
    bool busy = false;
};

// 32-bit ids -> 16-bit ids (8 per thread); any id that does not fit raises SWT_ERR_RANGE in the status word.
// The token count is read from the status words of the encode call that precedes it on the stream.
__global__ void __launch_bounds__(256) narrow_ids_kernel(const uint32_t *__restrict__ ids, uint16_t *__restrict__ out,
                                                         uint32_t *status) {
    const uint64_t n = ((uint64_t)status[swt::kStatusTokensHi] << 32) | status[swt::kStatusTokens];
    if (status[swt::kStatusCode] != SWT_OK) return;
    uint32_t wide = 0;
    for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += (uint64_t)gridDim.x * blockDim.x * 8) {
        if (i + 8 <= n) {
            const uint4 a = *reinterpret_cast<const uint4 *>(ids + i), b = *reinterpret_cast<const uint4 *>(ids + i + 4);
            wide |= a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w;
            *reinterpret_cast<uint4 *>(out + i) = make_uint4((a.x & 0xFFFFu) | (a.y << 16), (a.z & 0xFFFFu) | (a.w << 16),
                                                            (b.x & 0xFFFFu) | (b.y << 16), (b.z & 0xFFFFu) | (b.w << 16));
        } else {
            for (uint64_t k = i; k < n; ++k) { wide |= ids[k]; out[k] = (uint16_t)ids[k]; }
        }
    }
    if (wide > 0xFFFFu) atomicExch(&status[swt::kStatusCode], (uint32_t)SWT_ERR_RANGE);
}
}  // namespace

struct swt_pipeline {
    int device;
    uint64_t batch_bytes, max_words;
    Slot slot[kSlots];
};

using namespace swt;

SWT_API int swt_pipeline_create(int device, uint64_t batch_bytes, swt_pipeline **out) {
    SWT_REQUIRE(out != nullptr, "out is NULL");
    SWT_REQUIRE(batch_bytes >= (1u << 16) && batch_bytes <= (1ull << 30), "batch_bytes must be in [64 KiB, 1 GiB]");
    SWT_CUDA_OK(cudaSetDevice(device));
    swt_pipeline *p = new swt_pipeline();
    // Sized for the worst case of the BERT pre-tokenizer, where every punctuation character is a word of its own ("a,a,a,..."
    // or "....": one word per byte).  Everything a host call needs is allocated here: the hot calls allocate nothing.
    p->device = device; p->batch_bytes = batch_bytes; p->max_words = batch_bytes;
    for (int i = 0; i < kSlots; ++i) {
        Slot &s = p->slot[i];
        // BPE long-word scratch is sized for the worst case (every word of the batch is long)
        s.ws_bytes = swt_encode_workspace_bytes((uint32_t)p->max_words, batch_bytes * kLowerGrowthNum / kLowerGrowthDen);
        cudaError_t e = cudaMalloc(&s.d_arena, batch_bytes * kLowerGrowthNum / kLowerGrowthDen + 16);
        if (e == cudaSuccess) e = cudaMalloc(&s.d_off, (p->max_words + 1) * 4);
        if (e == cudaSuccess) e = cudaMalloc(&s.d_tok_off, (p->max_words + 1) * 4);
        if (e == cudaSuccess) e = cudaMalloc(&s.d_ids, (batch_bytes * kLowerGrowthNum / kLowerGrowthDen + p->max_words + 16) * 4);
        if (e == cudaSuccess) e = cudaMalloc(&s.d_ids16, (batch_bytes * kLowerGrowthNum / kLowerGrowthDen + p->max_words + 16) * 2);
        if (e == cudaSuccess) e = cudaMalloc(&s.d_text, batch_bytes + 16);
        s.ptws_bytes = swt_pretok_workspace_bytes(batch_bytes);
        if (e == cudaSuccess) e = cudaMalloc(&s.d_ptws, s.ptws_bytes);
        if (e == cudaSuccess) e = cudaMalloc(&s.d_status, 8 * 4);
        if (e == cudaSuccess) e = cudaMalloc(&s.d_ws, s.ws_bytes);
        if (e == cudaSuccess) e = cudaHostAlloc((void **)&s.h_status, 8 * 4, cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.kernel_done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.d2h_done, cudaEventDisableTiming);
        if (e != cudaSuccess) {
            set_error(std::string("pipeline allocation: ") + cudaGetErrorString(e));
            swt_pipeline_destroy(p);
            return SWT_ERR_CUDA;
        }
    }
    *out = p;
    return SWT_OK;
}

SWT_API void swt_pipeline_destroy(swt_pipeline *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    for (int i = 0; i < kSlots; ++i) {
        Slot &s = p->slot[i];
        if (s.stream) cudaStreamSynchronize(s.stream);
        cudaFree(s.d_arena); cudaFree(s.d_off); cudaFree(s.d_tok_off); cudaFree(s.d_ids); cudaFree(s.d_ids16); cudaFree(s.d_text); cudaFree(s.d_ptws); cudaFree(s.d_status); cudaFree(s.d_ws);
        if (s.h_status) cudaFreeHost(s.h_status);
        if (s.kernel_done) cudaEventDestroy(s.kernel_done);
        if (s.d2h_done) cudaEventDestroy(s.d2h_done);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    delete p;
}

static int encode_host_impl(swt_pipeline *p, int which, const void *table, const uint8_t *h_arena, const uint32_t *h_word_off,
                            uint64_t n_words, void *h_out_ids_any, bool narrow, uint64_t out_cap, uint32_t *h_out_tok_off,
                            uint64_t *n_tokens, uint64_t *h6_events) {
    uint32_t *h_out_ids = narrow ? nullptr : (uint32_t *)h_out_ids_any;
    uint16_t *h_out_ids16 = narrow ? (uint16_t *)h_out_ids_any : nullptr;
    SWT_REQUIRE(p && table && h_word_off && n_tokens, "NULL argument");
    SWT_REQUIRE(which == 0 || which == 1, "which must be 0 (BPE) or 1 (WP)");
    SWT_REQUIRE(n_words < 0xFFFFFFFFull, "n_words must be < 2^32 per call");
    SWT_REQUIRE(n_words == 0 || (h_arena && h_out_ids_any), "NULL data pointer");
    SWT_CUDA_OK(cudaSetDevice(p->device));
    // batch boundaries: [w0, w1) with at most batch_bytes bytes and max_words words
    struct Batch { uint64_t w0, w1; };
    std::vector<Batch> batches;
    for (uint64_t w0 = 0; w0 < n_words;) {
        const uint64_t limit = (uint64_t)h_word_off[w0] + p->batch_bytes;
        const uint64_t hi = std::min<uint64_t>(n_words, w0 + p->max_words);
        // largest w1 in (w0, hi] with off[w1] <= limit
        const uint32_t *it = std::upper_bound(h_word_off + w0 + 1, h_word_off + hi + 1, limit,
                                              [](uint64_t v, uint32_t o) { return v < (uint64_t)o; });
        uint64_t w1 = (uint64_t)(it - h_word_off) - 1;
        if (w1 <= w0) { set_error("a single word exceeds the pipeline batch size"); return SWT_ERR_CAPACITY; }
        batches.push_back({w0, w1});
        w0 = w1;
    }
    auto enqueue_h2d = [&](size_t k) -> int {
        Slot &s = p->slot[k % kSlots];
        if (s.busy) { SWT_CUDA_OK(cudaEventSynchronize(s.d2h_done)); s.busy = false; }
        const Batch &b = batches[k];
        const uint64_t byte0 = h_word_off[b.w0], nbytes = h_word_off[b.w1] - byte0;
        if (nbytes) SWT_CUDA_OK(cudaMemcpyAsync(s.d_arena, h_arena + byte0, nbytes, cudaMemcpyHostToDevice, s.stream));
        SWT_CUDA_OK(cudaMemcpyAsync(s.d_off, h_word_off + b.w0, (b.w1 - b.w0 + 1) * 4, cudaMemcpyHostToDevice, s.stream));
        return SWT_OK;
    };
    uint64_t total = 0, h6 = 0;
    int rc = SWT_OK;
    if (!batches.empty()) rc = enqueue_h2d(0);
    for (size_t k = 0; k < batches.size() && rc == SWT_OK; ++k) {
        if (k + 1 < batches.size()) { rc = enqueue_h2d(k + 1); if (rc) break; }
        Slot &s = p->slot[k % kSlots];
        const Batch &b = batches[k];
        const uint32_t nw = (uint32_t)(b.w1 - b.w0);
        const uint64_t byte0 = h_word_off[b.w0], nbytes = h_word_off[b.w1] - byte0;
        const uint64_t cap = nbytes + nw + 16;
        if (total > 0xFFFFFFFFull) { set_error("more than 2^32 tokens in one call"); rc = SWT_ERR_CAPACITY; break; }
        // offsets stay absolute: hand the kernel an arena pointer rebased by the batch's first byte
        const uint8_t *arena_rebased = s.d_arena - byte0;
        if (which == 0)
            rc = bpe_encode_launch((const swt_bpe_table *)table, arena_rebased, s.d_off, nw, nbytes, s.d_ids, cap,
                                   h_out_tok_off ? s.d_tok_off : nullptr, (uint32_t)total, s.d_ws, s.ws_bytes, s.d_status, s.stream);
        else
            rc = wp_encode_launch((const swt_wp_trie *)table, arena_rebased, s.d_off, nw, nbytes, s.d_ids, cap,
                                  h_out_tok_off ? s.d_tok_off : nullptr, (uint32_t)total, s.d_ws, s.ws_bytes, s.d_status, s.stream);
        if (rc) break;
        if (narrow) narrow_ids_kernel<<<swt::kNumSMs * 8, 256, 0, s.stream>>>(s.d_ids, s.d_ids16, s.d_status);
        SWT_CUDA_BRK(cudaMemcpyAsync(s.h_status, s.d_status, 8 * 4, cudaMemcpyDeviceToHost, s.stream));
        SWT_CUDA_BRK(cudaEventRecord(s.kernel_done, s.stream));
        SWT_CUDA_BRK(cudaEventSynchronize(s.kernel_done));
        if (s.h_status[kStatusCode] != SWT_OK) { set_error("encode kernel reported status " + std::to_string(s.h_status[kStatusCode])); rc = (int)s.h_status[kStatusCode]; break; }
        const uint64_t nt = ((uint64_t)s.h_status[kStatusTokensHi] << 32) | s.h_status[kStatusTokens];
        h6 += s.h_status[kStatusH6];
        if (total + nt > out_cap) { set_error("h_out_ids capacity too small"); rc = SWT_ERR_CAPACITY; break; }
        if (nt) SWT_CUDA_BRK(narrow ? cudaMemcpyAsync(h_out_ids16 + total, s.d_ids16, nt * 2, cudaMemcpyDeviceToHost, s.stream)
                                    : cudaMemcpyAsync(h_out_ids + total, s.d_ids, nt * 4, cudaMemcpyDeviceToHost, s.stream));
        if (h_out_tok_off) SWT_CUDA_BRK(cudaMemcpyAsync(h_out_tok_off + b.w0, s.d_tok_off, (uint64_t)nw * 4, cudaMemcpyDeviceToHost, s.stream));
        SWT_CUDA_BRK(cudaEventRecord(s.d2h_done, s.stream));
        s.busy = true;
        total += nt;
    }
    for (int i = 0; i < kSlots; ++i) { cudaStreamSynchronize(p->slot[i].stream); p->slot[i].busy = false; }
    if (rc != SWT_OK) return rc;
    if (h_out_tok_off) h_out_tok_off[n_words] = (uint32_t)total;
    *n_tokens = total;
    if (h6_events) *h6_events = h6;
    return SWT_OK;
}

SWT_API int swt_encode_host(swt_pipeline *p, int which, const void *table, const uint8_t *h_arena, const uint32_t *h_word_off,
                            uint64_t n_words, uint32_t *h_out_ids, uint64_t out_cap, uint32_t *h_out_tok_off,
                            uint64_t *n_tokens, uint64_t *h6_events) {
    return encode_host_impl(p, which, table, h_arena, h_word_off, n_words, h_out_ids, false, out_cap, h_out_tok_off, n_tokens, h6_events);
}

SWT_API int swt_encode_host16(swt_pipeline *p, int which, const void *table, const uint8_t *h_arena, const uint32_t *h_word_off,
                              uint64_t n_words, uint16_t *h_out_ids16, uint64_t out_cap, uint32_t *h_out_tok_off,
                              uint64_t *n_tokens, uint64_t *h6_events) {
    return encode_host_impl(p, which, table, h_arena, h_word_off, n_words, h_out_ids16, true, out_cap, h_out_tok_off, n_tokens, h6_events);
}

// ---- raw text in, flat token ids out: pre-tokenization (pretok.cu) + FastBPE / FastWP encode per batch --------------------------------
// batch cut points: Python's str.split treats 0x1C-0x1F as whitespace (FastWP); Rust's char::is_whitespace, which the BERT
// pre-tokenizer of the BPE classes uses, does not -- cutting there would split a word such as "ab\x1ccd"
static bool ascii_space(uint8_t b, bool bert) { return b == 0x20 || (b >= 0x09 && b <= 0x0D) || (!bert && b >= 0x1C && b <= 0x1F); }

SWT_API int swt_tokenize_text_host(swt_pipeline *p, const swt_pretok *pretok, int which, const void *table, const uint8_t *h_text,
                                   uint64_t n_bytes, void *h_out_ids, int ids_16bit, uint64_t out_cap, uint64_t *n_tokens,
                                   uint64_t *n_words_out, uint64_t *h6_events) {
    SWT_REQUIRE(p && pretok && table && n_tokens, "NULL argument");
    SWT_REQUIRE(which == 0 || which == 1, "which must be 0 (BPE) or 1 (WP)");
    SWT_REQUIRE(n_bytes == 0 || (h_text && h_out_ids), "NULL data pointer");
    SWT_CUDA_OK(cudaSetDevice(p->device));
    const bool narrow = ids_16bit != 0;
    const bool bert = swt::pretok_mode(pretok) == SWT_PRETOK_BERT;
    const uint64_t slice_max = p->batch_bytes - 8;
    // batches end just after an ASCII whitespace byte (a whole character in UTF-8), so no word is cut
    struct Batch { uint64_t b0, b1; };
    std::vector<Batch> batches;
    for (uint64_t b0 = 0; b0 < n_bytes;) {
        uint64_t b1 = std::min<uint64_t>(n_bytes, b0 + slice_max);
        if (b1 < n_bytes) {
            uint64_t j = b1;
            while (j > b0 && !ascii_space(h_text[j - 1], bert)) --j;
            if (j == b0) { set_error("no ASCII whitespace within one pipeline batch: raise batch_bytes"); return SWT_ERR_CAPACITY; }
            b1 = j;
        }
        batches.push_back({b0, b1});
        b0 = b1;
    }
    auto enqueue_h2d = [&](size_t k) -> int {
        Slot &s = p->slot[k % kSlots];
        if (s.busy) { SWT_CUDA_OK(cudaEventSynchronize(s.d2h_done)); s.busy = false; }
        const Batch &b = batches[k];
        const uint64_t nb = b.b1 - b.b0;
        SWT_CUDA_OK(cudaMemcpyAsync(s.d_text, h_text + b.b0, nb, cudaMemcpyHostToDevice, s.stream));
        SWT_CUDA_OK(cudaMemsetAsync(s.d_text + nb, 0, 8, s.stream));
        int rc = swt_pretok_count(pretok, s.d_text, nb, s.d_ptws, s.ptws_bytes, s.d_status, s.stream);
        if (rc) return rc;
        SWT_CUDA_OK(cudaMemcpyAsync(s.h_status, s.d_status, 8 * 4, cudaMemcpyDeviceToHost, s.stream));
        SWT_CUDA_OK(cudaEventRecord(s.kernel_done, s.stream));
        return SWT_OK;
    };
    uint64_t total = 0, h6 = 0, words = 0;
    int rc = SWT_OK;
    if (!batches.empty()) rc = enqueue_h2d(0);
    for (size_t k = 0; k < batches.size() && rc == SWT_OK; ++k) {
        if (k + 1 < batches.size()) { rc = enqueue_h2d(k + 1); if (rc) break; }
        Slot &s = p->slot[k % kSlots];
        const uint64_t nb = batches[k].b1 - batches[k].b0;
        SWT_CUDA_BRK(cudaEventSynchronize(s.kernel_done));                       // counts of this batch
        if (s.h_status[0] != SWT_OK) { set_error("pre-tokenizer reported status " + std::to_string(s.h_status[0])); rc = (int)s.h_status[0]; break; }
        const uint32_t nw = s.h_status[1];
        const uint64_t n_arena = ((uint64_t)s.h_status[3] << 32) | s.h_status[2];
        if (nw > p->max_words) { set_error("more words in a batch than the pipeline was sized for"); rc = SWT_ERR_CAPACITY; break; }
        rc = swt_pretok_write(pretok, s.d_text, nb, s.d_ptws, s.ptws_bytes, s.d_arena, n_arena, s.d_off, nullptr, p->max_words + 1, nw, n_arena,
                              s.d_status, s.stream);
        if (rc) break;
        words += nw;
        if (nw == 0) continue;
        if (which == 0)
            rc = bpe_encode_launch((const swt_bpe_table *)table, s.d_arena, s.d_off, nw, n_arena, s.d_ids, n_arena + nw + 16, nullptr, 0, s.d_ws,
                                   s.ws_bytes, s.d_status, s.stream);
        else
            rc = wp_encode_launch((const swt_wp_trie *)table, s.d_arena, s.d_off, nw, n_arena, s.d_ids, n_arena + nw + 16, nullptr, 0, s.d_ws, s.ws_bytes,
                                  s.d_status, s.stream);
        if (rc) break;
        if (narrow) narrow_ids_kernel<<<swt::kNumSMs * 8, 256, 0, s.stream>>>(s.d_ids, s.d_ids16, s.d_status);
        SWT_CUDA_BRK(cudaMemcpyAsync(s.h_status, s.d_status, 8 * 4, cudaMemcpyDeviceToHost, s.stream));
        SWT_CUDA_BRK(cudaEventRecord(s.kernel_done, s.stream));
        SWT_CUDA_BRK(cudaEventSynchronize(s.kernel_done));
        if (s.h_status[kStatusCode] != SWT_OK) { set_error("encode kernel reported status " + std::to_string(s.h_status[kStatusCode])); rc = (int)s.h_status[kStatusCode]; break; }
        const uint64_t nt = ((uint64_t)s.h_status[kStatusTokensHi] << 32) | s.h_status[kStatusTokens];
        h6 += s.h_status[kStatusH6];
        if (total + nt > out_cap) { set_error("h_out_ids capacity too small"); rc = SWT_ERR_CAPACITY; break; }
        if (nt) SWT_CUDA_BRK(narrow ? cudaMemcpyAsync((uint16_t *)h_out_ids + total, s.d_ids16, nt * 2, cudaMemcpyDeviceToHost, s.stream)
                                    : cudaMemcpyAsync((uint32_t *)h_out_ids + total, s.d_ids, nt * 4, cudaMemcpyDeviceToHost, s.stream));
        SWT_CUDA_BRK(cudaEventRecord(s.d2h_done, s.stream));
        s.busy = true;
        total += nt;
    }
    for (int i = 0; i < kSlots; ++i) { cudaStreamSynchronize(p->slot[i].stream); p->slot[i].busy = false; }
    if (rc != SWT_OK) return rc;
    *n_tokens = total;
    if (n_words_out) *n_words_out = words;
    if (h6_events) *h6_events = h6;
    return SWT_OK;
}

// ---- small calls: one short text per call (the per-line tokenize() pattern of the reference's CLI, cli.py:253-264) ----------------------
// Everything the call needs is owned by the swt_small object: a pinned, device-mapped input and output buffer (zero copy: the
// kernels read the text from and write the ids to host memory), device scratch, one stream.  A call is a memcpy of the text into
// the pinned buffer, ONE single-CTA launch (pre-tokenizer + encoder, tokenize_small_kernel) and ONE synchronisation; nothing is allocated.
struct swt_small {
    int device;
    uint32_t max_bytes, out_cap;
    uint8_t *h_in = nullptr, *d_in = nullptr;        // pinned + mapped: host and device view of the text
    uint32_t *h_out = nullptr, *d_out = nullptr;     // pinned + mapped: header (8 words) + ids
    uint8_t *d_blob = nullptr;                       // device scratch
    uint8_t *d_arena; uint32_t *d_off, *d_scratch, *d_cnt, *d_compact, *d_long;
    cudaStream_t stream = nullptr;
    uint32_t seq = 0;                                // sequence number of the last call (completion flag)
    struct Bound { const swt_pretok *pretok; const void *table; int which, naive; };
    static constexpr uint32_t kMaxBound = 64;
    Bound bound[kMaxBound]; uint32_t n_bound = 0;
};

SWT_API int swt_small_create(int device, swt_small **out) {
    SWT_REQUIRE(out != nullptr, "out is NULL");
    SWT_CUDA_OK(cudaSetDevice(device));
    swt_small *s = new swt_small();
    s->device = device;
    s->max_bytes = pretok_small_max_bytes();
    const size_t arena_cap = (size_t)s->max_bytes * kLowerGrowthNum / kLowerGrowthDen + 64, max_words = s->max_bytes + 1;
    s->out_cap = (uint32_t)(arena_cap + max_words);
    Carver sz(nullptr);
    sz.take<uint8_t>(arena_cap); sz.take<uint32_t>(max_words + 1); sz.take<uint32_t>(s->out_cap + 16); sz.take<uint32_t>(max_words);
    sz.take<uint32_t>(s->out_cap + 16); sz.take<uint32_t>(2 * arena_cap + 32);
    cudaError_t e = cudaHostAlloc((void **)&s->h_in, s->max_bytes + 64, cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&s->h_out, ((size_t)s->out_cap + 16) * 4, cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void **)&s->d_in, s->h_in, 0);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void **)&s->d_out, s->h_out, 0);
    if (e == cudaSuccess) e = cudaMalloc((void **)&s->d_blob, sz.used());
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { set_error(std::string("swt_small_create: ") + cudaGetErrorString(e)); swt_small_destroy(s); return SWT_ERR_CUDA; }
    Carver cv(s->d_blob);
    s->d_arena = cv.take<uint8_t>(arena_cap); s->d_off = cv.take<uint32_t>(max_words + 1);
    s->d_scratch = cv.take<uint32_t>(s->out_cap + 16); s->d_cnt = cv.take<uint32_t>(max_words); s->d_compact = cv.take<uint32_t>(s->out_cap + 16);
    s->d_long = cv.take<uint32_t>(2 * arena_cap + 32);
    memset(s->h_out, 0, 64);
    *out = s;
    return SWT_OK;
}

SWT_API void swt_small_destroy(swt_small *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) { cudaStreamSynchronize(s->stream); cudaStreamDestroy(s->stream); }
    cudaFree(s->d_blob);
    if (s->h_in) cudaFreeHost(s->h_in);
    if (s->h_out) cudaFreeHost(s->h_out);
    delete s;
}

SWT_API uint32_t swt_small_max_bytes(void) { return pretok_small_max_bytes(); }

// One call: memcpy of the text into the pinned buffer, one launch, then the host polls the completion flag the kernel writes last
// (out[4] = sequence number of the call; cheaper than a stream synchronisation); a stream query every few thousand polls catches a
// failed launch.
static int small_run(swt_small *s, const swt_pretok *pretok, int which, const void *table, int naive, const uint8_t *text, uint32_t n_bytes) {
    s->h_out[1] = 0;
    if (n_bytes == 0) { s->h_out[0] = SWT_OK; s->h_out[2] = s->h_out[3] = 0; return SWT_OK; }
    if (n_bytes > kInlineTextBytes) {                  // shorter texts ride in the kernel's parameter buffer
        memcpy(s->h_in, text, n_bytes);
        memset(s->h_in + n_bytes, 0, 16);
    }
    const uint32_t seq = ++s->seq;
    const SmallArgs a{s->d_arena, s->d_off, s->d_scratch, s->d_cnt, s->d_compact, s->d_long, s->d_out, s->out_cap, seq};
    const bool bert = swt::pretok_mode(pretok) == SWT_PRETOK_BERT;
    const int rc = which == 0 ? bpe_small_launch((const swt_bpe_table *)table, naive, pretok_dev_view(pretok), bert, text, s->d_in, n_bytes, a, s->stream)
                              : wp_small_launch((const swt_wp_trie *)table, naive, pretok_dev_view(pretok), bert, text, s->d_in, n_bytes, a, s->stream);
    if (rc) return rc;
    volatile uint32_t *flag = s->h_out + 4;
    for (uint32_t spins = 0; *flag != seq; ++spins) {
        if ((spins & 0xFFFu) == 0xFFFu) {
            const cudaError_t e = cudaStreamQuery(s->stream);
            if (e == cudaSuccess) break;                                             // finished: the flag is (about to be) visible
            if (e != cudaErrorNotReady) { set_error(std::string("small-call kernel: ") + cudaGetErrorString(e)); return SWT_ERR_CUDA; }
        }
    }
    if (*flag != seq) SWT_CUDA_OK(cudaStreamSynchronize(s->stream));
    if (s->h_out[0] != SWT_OK) { set_error("small-call kernel reported status " + std::to_string(s->h_out[0])); return (int)s->h_out[0]; }
    return SWT_OK;
}

SWT_API int swt_tokenize_small(swt_small *s, const swt_pretok *pretok, int which, const void *table, int naive, const uint8_t *text, uint32_t n_bytes,
                               const uint32_t **ids, uint32_t *n_tokens, uint32_t *n_words, uint32_t *h6_events) {
    SWT_REQUIRE(s && pretok && table && ids && n_tokens, "NULL argument");
    SWT_REQUIRE(which == 0 || which == 1, "which must be 0 (BPE) or 1 (WP)");
    SWT_REQUIRE(n_bytes <= s->max_bytes, "text longer than swt_small_max_bytes()");
    SWT_REQUIRE(n_bytes == 0 || text, "text is NULL");
    *ids = s->h_out + 8; *n_tokens = 0;
    if (n_words) *n_words = 0;
    if (h6_events) *h6_events = 0;
    const int rc = small_run(s, pretok, which, table, naive, text, n_bytes);
    if (rc) return rc;
    *n_tokens = s->h_out[1];
    if (h6_events) *h6_events = s->h_out[2];
    if (n_words) *n_words = s->h_out[3];
    return SWT_OK;
}

// The same call for hosts where every argument conversion counts (ctypes): the (pre-tokenizer, table) pair is bound once, the call
// takes the text only and the results are read from the output buffer (swt_small_output: [0] status, [1] tokens, [2] H6 events,
// [3] words, ids from word 8).
SWT_API int swt_small_bind(swt_small *s, const swt_pretok *pretok, int which, const void *table, int naive, uint32_t *binding) {
    SWT_REQUIRE(s && pretok && table && binding, "NULL argument");
    SWT_REQUIRE(which == 0 || which == 1, "which must be 0 (BPE) or 1 (WP)");
    for (uint32_t k = 0; k < s->n_bound; ++k)
        if (s->bound[k].pretok == pretok && s->bound[k].table == table && s->bound[k].which == which && s->bound[k].naive == naive) { *binding = k; return SWT_OK; }
    uint32_t k = 0;
    while (k < s->n_bound && s->bound[k].table != nullptr) ++k;                     // first free slot
    if (k == swt_small::kMaxBound) { set_error("swt_small_bind: all bindings in use (swt_small_unbind the ones of destroyed tables)"); return SWT_ERR_CAPACITY; }
    if (k == s->n_bound) ++s->n_bound;
    s->bound[k] = {pretok, table, which, naive};
    *binding = k;
    return SWT_OK;
}
SWT_API void swt_small_unbind(swt_small *s, uint32_t binding) {
    if (s && binding < s->n_bound) s->bound[binding] = {nullptr, nullptr, 0, 0};
}
SWT_API const uint32_t *swt_small_output(const swt_small *s) { return s ? s->h_out : nullptr; }
SWT_API int swt_tokenize_small_bound(swt_small *s, uint32_t binding, const uint8_t *text, uint32_t n_bytes) {
    SWT_REQUIRE(s && binding < s->n_bound && s->bound[binding].table != nullptr, "unknown binding");
    SWT_REQUIRE(n_bytes <= s->max_bytes, "text longer than swt_small_max_bytes()");
    const auto &b = s->bound[binding];
    return small_run(s, b.pretok, b.which, b.table, b.naive, text, n_bytes);
}
