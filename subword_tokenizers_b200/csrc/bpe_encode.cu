// bpe_encode.cu -- HP-1: FastBPE.encode_word (reference source/bpe.py:205-243) on sm_100a.
//
// Rank table: open-addressing hash, 16-byte slots {left, right, rank, merged}; one 128-bit load per
// probe, key = left<<32|right.  Built on the host from the merge list (last rank wins for a pair that
// is listed twice, bpe.py:200,257) and uploaded once per table.
//
// Kernel: see encode.cuh for the tile kernel and the word-type memo.  Per directly encoded word (one thread,
// symbols in a thread-local buffer): repeat { min rank over the adjacent pairs (bpe.py:212-217); greedy left-to-right
// replacement of every occurrence (:221-235) } until no ranked pair is left or one symbol remains.
// Words longer than kShortBytes are processed by a whole warp in global scratch.
#include <vector>

#include "encode.cuh"

namespace swt {

struct BpeTableDev {
    const uint4 *slots;      // {left, right, rank, merged}; rank == 0xFFFFFFFF marks an empty slot
    uint32_t mask;           // n_slots - 1
    const uint32_t *bmp_lut; // 65536 entries: code point -> symbol id or 0xFFFFFFFF
    const uint32_t *hi_cp;   // sorted code points >= 0x10000 that have a symbol
    const uint32_t *hi_id;
    uint32_t n_hi;
    const uint32_t *m_left, *m_right, *m_new;   // by rank
    uint32_t n_merges;
};

}  // namespace swt

struct swt_bpe_table {
    swt::BpeTableDev dev;
    int device;
    void *d_blob;
};

namespace swt {

constexpr uint32_t kEmptyRank = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t bpe_char_symbol(const BpeTableDev &t, uint32_t cp) {
    uint32_t id = 0xFFFFFFFFu;
    if (cp < 0x10000u) id = __ldg(&t.bmp_lut[cp]);
    else {
        uint32_t lo = 0, hi = t.n_hi;
        while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (__ldg(&t.hi_cp[mid]) < cp) lo = mid + 1; else hi = mid; }
        if (lo < t.n_hi && __ldg(&t.hi_cp[lo]) == cp) id = __ldg(&t.hi_id[lo]);
    }
    return id == 0xFFFFFFFFu ? (SWT_BPE_UNKNOWN_CP | cp) : id;
}

// returns rank (kEmptyRank when the pair is not in the table); merged id in `merged`
__device__ __forceinline__ uint32_t bpe_probe(const BpeTableDev &t, uint32_t a, uint32_t b, uint32_t &merged) {
    if ((a | b) & SWT_BPE_UNKNOWN_CP) return kEmptyRank;          // characters no merge mentions
    uint32_t h = (uint32_t)mix64(((uint64_t)a << 32) | b) & t.mask;
    for (;;) {
        uint4 e = __ldg(&t.slots[h]);
        if (e.z == kEmptyRank) return kEmptyRank;
        if (e.x == a && e.y == b) { merged = e.w; return e.z; }
        h = (h + 1) & t.mask;
    }
}

// ---- short words: one thread, symbols in a thread-local buffer ----------------------------------------------------
// kNaive = NaiveBPE.encode_word (bpe.py:114-132): the merges are replayed in list order, i.e. the next merge applied is the
// lowest-ranked pair present whose rank is ABOVE the last one applied (a pair that only appears after its turn is skipped).
template <bool kNaive>
__device__ __forceinline__ uint32_t bpe_encode_short(const BpeTableDev &t, const uint8_t *p, uint32_t nbytes, uint32_t *s) {
    uint32_t n = 0;
    for (uint32_t i = 0; i < nbytes;) {
        uint32_t adv; uint32_t cp = utf8_decode(p + i, nbytes - i, adv); i += adv;
        s[n++] = bpe_char_symbol(t, cp);
    }
    if (n == 0) { if (kNaive) return 0; s[0] = SWT_BPE_EMPTY_TOKEN; return 1; }   // FastBPE: [""] (bpe.py:207-208); NaiveBPE: [] (:131-132)
    uint32_t last = kEmptyRank;                                     // kNaive: rank of the merge applied last
    while (n >= 2) {
        uint32_t best = kEmptyRank, ba = 0, bb = 0, bz = 0;
        uint32_t prev = s[0];
        for (uint32_t i = 0; i + 1 < n; ++i) {                      // min rank over the adjacent pairs (bpe.py:212-217)
            uint32_t cur = s[i + 1], z;
            uint32_t r = bpe_probe(t, prev, cur, z);
            if (r < best && (!kNaive || last == kEmptyRank || r > last)) { best = r; ba = prev; bb = cur; bz = z; }
            prev = cur;
        }
        if (best == kEmptyRank) break;
        last = best;
        uint32_t r = 0, o = 0;                                      // greedy left-to-right replacement (bpe.py:221-235)
        while (r < n) {
            uint32_t v = s[r];
            if (r + 1 < n && v == ba && s[r + 1] == bb) { s[o] = bz; r += 2; }
            else { s[o] = v; r += 1; }
            ++o;
        }
        n = o;
    }
    for (uint32_t k = 0; k < n; ++k) s[k] = (s[k] << 1) | (k > 0);   // '##' prefix of symbols[1:] (bpe.py:240-241)
    return n;
}

// ---- long words: one WARP works on the word in global scratch ----------------------------------------------------------
// bufA/bufB are ping-pong symbol buffers of at least nbytes entries.  Returns the final symbol count; *result points
// at the buffer holding the final symbols (already in token form).  All 32 lanes must call this.
template <bool kNaive>
__device__ __noinline__ uint32_t bpe_encode_long_warp(const BpeTableDev &t, const uint8_t *p, uint32_t nbytes, uint32_t *bufA,
                                                      uint32_t *bufB, uint32_t **result) {
    const uint32_t lane = threadIdx.x & 31;
    auto warp_excl_scan = [&](uint32_t v, uint32_t &total) {
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += u; }
        total = __shfl_sync(0xffffffffu, incl, 31);
        return incl - v;
    };
    // 1. UTF-8 decode, 32 bytes at a time: a byte starts a character unless it is 10xxxxxx
    uint32_t n = 0;
    for (uint32_t base = 0; base < nbytes; base += 32) {
        const uint32_t i = base + lane;
        const uint32_t is_start = (i < nbytes) && ((p[i] & 0xC0u) != 0x80u);
        uint32_t total; const uint32_t excl = warp_excl_scan(is_start, total);
        if (is_start) { uint32_t adv; const uint32_t cp = utf8_decode(p + i, nbytes - i, adv); bufA[n + excl] = bpe_char_symbol(t, cp); }
        n += total;
    }
    __syncwarp();
    uint32_t *src = bufA, *dst = bufB;
    uint32_t last = kEmptyRank;
    while (n >= 2) {
        // 2. min rank over all adjacent pairs (bpe.py:212-217)
        uint32_t best = kEmptyRank;
        for (uint32_t i = lane; i + 1 < n; i += 32) {
            uint32_t z; const uint32_t r = bpe_probe(t, src[i], src[i + 1], z);
            if (!kNaive || last == kEmptyRank || r > last) best = min(best, r);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, d));
        if (best == kEmptyRank) break;
        last = best;
        const uint32_t a = __ldg(&t.m_left[best]), b = __ldg(&t.m_right[best]), z = __ldg(&t.m_new[best]);
        // 3. greedy left-to-right replacement (bpe.py:221-235); each lane owns one contiguous segment
        const uint32_t seg = (n + 31) / 32;
        const uint32_t lo = min(n, lane * seg), hi = min(n, lo + seg);
        // is element `lo` the right half of a pair selected by an earlier segment?  For a != b matches cannot overlap;
        // for a == b walk back over the run of a's to get the parity.
        uint32_t i = lo;
        if (lo < hi && lo > 0) {
            if (a != b) { if (src[lo - 1] == a && src[lo] == b) i = lo + 1; }
            else if (src[lo] == a) {
                uint32_t k = lo; while (k > 0 && src[k - 1] == a) --k;      // run start
                if ((lo - k) & 1u) i = lo + 1;                                // lo is consumed as a right half
            }
        }
        const uint32_t first = i;
        uint32_t cnt = 0;
        while (i < hi) { if (i + 1 < n && src[i] == a && src[i + 1] == b) i += 2; else i += 1; ++cnt; }
        uint32_t total; const uint32_t excl = warp_excl_scan(cnt, total);
        i = first; uint32_t o = excl;
        while (i < hi) {
            if (i + 1 < n && src[i] == a && src[i + 1] == b) { dst[o++] = z; i += 2; }
            else { dst[o++] = src[i]; i += 1; }
        }
        __syncwarp();
        n = total;
        uint32_t *tmp = src; src = dst; dst = tmp;
    }
    for (uint32_t k = lane; k < n; k += 32) src[k] = (src[k] << 1) | (k > 0);
    __syncwarp();
    *result = src;
    return n;
}

template <bool kNaive>
struct BpeEncT {
    BpeTableDev t;
    static constexpr bool kScratchLong = true;
    static constexpr bool kBatchSlowPath = false;
    __device__ __forceinline__ uint32_t encode_short(const uint8_t *p, uint32_t nbytes, uint32_t *buf, uint32_t &h6) const {
        (void)h6;
        return bpe_encode_short<kNaive>(t, p, nbytes, buf);
    }
    __device__ __forceinline__ uint32_t encode_long_warp(const uint8_t *p, uint32_t nbytes, uint32_t *bufA, uint32_t *bufB,
                                                         uint32_t **result) const {
        return bpe_encode_long_warp<kNaive>(t, p, nbytes, bufA, bufB, result);
    }
};
using BpeEnc = BpeEncT<false>;          // FastBPE.encode_word
using NaiveBpeEnc = BpeEncT<true>;      // NaiveBPE.encode_word (merge lists without repeated pairs)

}  // namespace swt

using namespace swt;

SWT_API int swt_bpe_table_create(const uint32_t *h_left, const uint32_t *h_right, const uint32_t *h_merged, uint32_t n_merges,
                                 const uint32_t *h_char_cp, const uint32_t *h_char_id, uint32_t n_chars, int device,
                                 swt_bpe_table **out) {
    SWT_REQUIRE(out != nullptr, "out is NULL");
    SWT_REQUIRE(n_merges == 0 || (h_left && h_right && h_merged), "merge arrays are NULL");
    SWT_REQUIRE(n_chars == 0 || (h_char_cp && h_char_id), "char arrays are NULL");
    SWT_REQUIRE(n_merges < 0x7FFFFFFFu, "too many merges");
    SWT_CUDA_OK(cudaSetDevice(device));
    const uint64_t n_slots = next_pow2((uint64_t)n_merges * 2 + 16);
    std::vector<uint4> slots(n_slots, make_uint4(0, 0, kEmptyRank, 0));
    for (uint32_t k = 0; k < n_merges; ++k) {
        SWT_REQUIRE(!((h_left[k] | h_right[k] | h_merged[k]) & 0xC0000000u), "symbol ids must be < 2^30");
        uint64_t h = mix64(((uint64_t)h_left[k] << 32) | h_right[k]) & (n_slots - 1);
        while (slots[h].z != kEmptyRank && !(slots[h].x == h_left[k] && slots[h].y == h_right[k])) h = (h + 1) & (n_slots - 1);
        slots[h] = make_uint4(h_left[k], h_right[k], k, h_merged[k]);      // last rank wins (dict semantics)
    }
    std::vector<uint32_t> lut(65536, 0xFFFFFFFFu), hi_cp, hi_id;
    for (uint32_t i = 0; i < n_chars; ++i) {
        SWT_REQUIRE(i == 0 || h_char_cp[i] > h_char_cp[i - 1], "char_cp must be strictly ascending");
        if (h_char_cp[i] < 0x10000u) lut[h_char_cp[i]] = h_char_id[i];
        else { hi_cp.push_back(h_char_cp[i]); hi_id.push_back(h_char_id[i]); }
    }
    // one device blob: slots | lut | hi_cp | hi_id | m_left | m_right | m_new
    Carver sz(nullptr);
    sz.take<uint4>(n_slots); sz.take<uint32_t>(65536); sz.take<uint32_t>(hi_cp.size() + 1); sz.take<uint32_t>(hi_id.size() + 1);
    sz.take<uint32_t>(n_merges + 1); sz.take<uint32_t>(n_merges + 1); sz.take<uint32_t>(n_merges + 1);
    void *blob = nullptr;
    SWT_CUDA_OK(cudaMalloc(&blob, sz.used()));
    Carver cv(blob);
    uint4 *d_slots = cv.take<uint4>(n_slots);
    uint32_t *d_lut = cv.take<uint32_t>(65536);
    uint32_t *d_hicp = cv.take<uint32_t>(hi_cp.size() + 1), *d_hiid = cv.take<uint32_t>(hi_id.size() + 1);
    uint32_t *d_l = cv.take<uint32_t>(n_merges + 1), *d_r = cv.take<uint32_t>(n_merges + 1), *d_n = cv.take<uint32_t>(n_merges + 1);
    cudaError_t e = cudaMemcpy(d_slots, slots.data(), n_slots * sizeof(uint4), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_lut, lut.data(), 65536 * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !hi_cp.empty()) e = cudaMemcpy(d_hicp, hi_cp.data(), hi_cp.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !hi_id.empty()) e = cudaMemcpy(d_hiid, hi_id.data(), hi_id.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n_merges) e = cudaMemcpy(d_l, h_left, n_merges * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n_merges) e = cudaMemcpy(d_r, h_right, n_merges * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n_merges) e = cudaMemcpy(d_n, h_merged, n_merges * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(blob); set_error(std::string("table upload: ") + cudaGetErrorString(e)); return SWT_ERR_CUDA; }
    swt_bpe_table *t = new swt_bpe_table();
    t->dev = BpeTableDev{d_slots, (uint32_t)(n_slots - 1), d_lut, d_hicp, d_hiid, (uint32_t)hi_cp.size(), d_l, d_r, d_n, n_merges};
    t->device = device; t->d_blob = blob;
    *out = t;
    return SWT_OK;
}

SWT_API void swt_bpe_table_destroy(swt_bpe_table *t) {
    if (!t) return;
    cudaSetDevice(t->device);
    cudaFree(t->d_blob);
    delete t;
}

namespace swt {
int encode_grid(const void *kernel, int block, size_t dyn_smem) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, dyn_smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    int dev = 0, sms = kNumSMs;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms * per_sm;
}

int bpe_encode_launch(const swt_bpe_table *t, const uint8_t *d_arena, const uint32_t *d_word_off, uint32_t n_words,
                      uint64_t long_word_bytes, uint32_t *d_out_ids, uint64_t out_cap, uint32_t *d_out_tok_off, uint32_t tok_base,
                      void *d_workspace, size_t workspace_bytes, uint32_t *d_status, cudaStream_t st) {
    SWT_REQUIRE(t != nullptr, "NULL table");
    return launch_encode_tiles(BpeEnc{t->dev}, d_arena, d_word_off, n_words, long_word_bytes, d_out_ids, out_cap, d_out_tok_off,
                               tok_base, d_workspace, workspace_bytes, d_status, st);
}
}  // namespace swt

SWT_API int swt_bpe_encode_naive(const swt_bpe_table *t, const uint8_t *d_arena, const uint32_t *d_word_off, uint32_t n_words,
                                 uint64_t long_word_bytes, uint32_t *d_out_ids, uint64_t out_cap, uint32_t *d_out_tok_off,
                                 void *d_workspace, size_t workspace_bytes, uint32_t *d_status, void *stream) {
    SWT_REQUIRE(t != nullptr, "NULL table");
    return launch_encode_tiles(NaiveBpeEnc{t->dev}, d_arena, d_word_off, n_words, long_word_bytes, d_out_ids, out_cap, d_out_tok_off,
                               0u, d_workspace, workspace_bytes, d_status, (cudaStream_t)stream);
}

SWT_API int swt_bpe_encode(const swt_bpe_table *t, const uint8_t *d_arena, const uint32_t *d_word_off, uint32_t n_words,
                           uint64_t long_word_bytes, uint32_t *d_out_ids, uint64_t out_cap, uint32_t *d_out_tok_off,
                           void *d_workspace, size_t workspace_bytes, uint32_t *d_status, void *stream) {
    return bpe_encode_launch(t, d_arena, d_word_off, n_words, long_word_bytes, d_out_ids, out_cap, d_out_tok_off, 0u,
                             d_workspace, workspace_bytes, d_status, (cudaStream_t)stream);
}
