// bpe_encode.cu -- HP-1: FastBPE.encode_word (reference source/bpe.py:205-243) on sm_100a.
//
// Rank table: open-addressing hash, 16-byte slots {left, right, rank, merged}; one 128-bit load per
// probe, key = left<<32|right.  Built on the host from the merge list (last rank wins for a pair that
// is listed twice, bpe.py:200,257) and uploaded once per table.
//
// Kernel: see encode.cuh for the tile kernel and the word-type memo.  Per directly encoded word:
// repeat { min rank over the adjacent pairs (bpe.py:212-217); greedy left-to-right replacement of every occurrence
// (:221-235) } until no ranked pair is left or one symbol remains.  Three shapes of that loop:
//   warp per word   (words <= 32 bytes, the few memo misses of a tile): lane i holds symbol i and probes the pair (i, i+1),
//                   the minimum rank is a shuffle reduction, the replacement a ballot + compaction shuffle.  The latency of
//                   a word is its number of merges (about 3), not its number of pair probes (about 12).
//   thread per word (many misses in one tile: one word per lane, symbols in a thread-local buffer)
//   warp per long word (> 32 bytes): symbols in global scratch, ping-pong buffers.
// Rank table in shared memory (north star): the full table (16 B x 2 x merges) does not fit, so every CTA stages
// (a) a Bloom filter over ALL ranked pairs -- most probes of the merge loop ask for pairs that are not ranked at all, and a
// filter miss answers them exactly without leaving the SM -- and (b) the lowest-rank merges, the ones applied most often, in a
// direct-mapped table; only filter hits that are not among those go to the L2-resident table.
#include <algorithm>
#include <cstring>
#include <vector>

#include "encode.cuh"

namespace swt {

constexpr uint32_t kBloomWords = 4096;   // 128 Kbit Bloom filter, two probes per pair
constexpr uint32_t kHotSlots = 256;      // direct-mapped table of the lowest-rank merges
struct BpeStage { uint32_t bloom[kBloomWords]; uint4 hot[kHotSlots]; };      // 20 KB of shared memory per CTA

struct BpeTableDev {
    const BpeStage *stage;   // global image of the shared-memory tables
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

// returns rank (kEmptyRank when the pair is not in the table); merged id in `merged`.  sg: the CTA's shared-memory tables or nullptr.
__device__ __forceinline__ uint32_t bpe_probe(const BpeTableDev &t, const BpeStage *sg, uint32_t a, uint32_t b, uint32_t &merged) {
    if ((a | b) & SWT_BPE_UNKNOWN_CP) return kEmptyRank;          // characters no merge mentions
    const uint64_t hh = mix64(((uint64_t)a << 32) | b);
    if (sg) {
        const uint32_t b1 = (uint32_t)(hh >> 32) & (kBloomWords * 32 - 1), b2 = (uint32_t)(hh >> 47) & (kBloomWords * 32 - 1);
        if (!((sg->bloom[b1 >> 5] >> (b1 & 31u)) & (sg->bloom[b2 >> 5] >> (b2 & 31u)) & 1u)) return kEmptyRank;   // exact negative
        const uint4 e = sg->hot[(uint32_t)(hh >> 24) & (kHotSlots - 1)];
        if (e.x == a && e.y == b && e.z != kEmptyRank) { merged = e.w; return e.z; }
    }
    uint32_t h = (uint32_t)hh & t.mask;
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
__device__ __forceinline__ uint32_t bpe_encode_short(const BpeTableDev &t, const BpeStage *sg, const uint8_t *p, uint32_t nbytes, uint32_t *s) {
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
            uint32_t r = bpe_probe(t, sg, prev, cur, z);
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

// ---- short words, warp per word (north-star item 2): lane i holds symbol i, probes the pair (i, i + 1); the minimum rank is a
// shuffle reduction; the greedy replacement is a ballot (runs of a == b resolved left to right) and a compaction shuffle.
// All 32 lanes must call this; `ids` (32 words of shared memory) receives the tokens.  Returns the token count.
template <bool kNaive>
__device__ __noinline__ uint32_t bpe_encode_short_warp(const BpeTableDev &t, const BpeStage *sg, const uint8_t *p, uint32_t nbytes, uint32_t *ids) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t byte = lane < nbytes ? p[lane] : 0x80u;
    const uint32_t smask = __ballot_sync(0xffffffffu, lane < nbytes && (byte & 0xC0u) != 0x80u);
    uint32_t n = __popc(smask);
    if (n == 0) { if (kNaive) return 0; if (lane == 0) ids[0] = SWT_BPE_EMPTY_TOKEN; __syncwarp(); return 1; }
    uint32_t sym = 0;
    if (lane < n) {
        const uint32_t pos = __fns(smask, 0, lane + 1);
        uint32_t adv; const uint32_t cp = utf8_decode(p + pos, nbytes - pos, adv);
        sym = bpe_char_symbol(t, cp);
    }
    uint32_t last = kEmptyRank;
    while (n >= 2) {
        const uint32_t nxt = __shfl_down_sync(0xffffffffu, sym, 1);
        const bool valid = lane + 1 < n;
        uint32_t z = 0, r = kEmptyRank;
        if (valid) r = bpe_probe(t, sg, sym, nxt, z);
        if (kNaive && last != kEmptyRank && r <= last) r = kEmptyRank;
        uint32_t best = r;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, d));
        if (best == kEmptyRank) break;
        last = best;
        const uint32_t wl = __ffs(__ballot_sync(0xffffffffu, r == best)) - 1;
        const uint32_t a = __shfl_sync(0xffffffffu, sym, wl), b = __shfl_sync(0xffffffffu, nxt, wl), zz = __shfl_sync(0xffffffffu, z, wl);
        const uint32_t mm = __ballot_sync(0xffffffffu, valid && sym == a && nxt == b);
        uint32_t sel = mm;
        if (a == b) {                                   // "aaa" -> [aa, a]: take a match, skip the one overlapping it (bpe.py:224-234)
            sel = 0;
            for (uint32_t x = mm; x;) { const uint32_t i = __ffs(x) - 1; sel |= 1u << i; x &= ~(3u << i); }
        }
        const uint32_t keep = ~(sel << 1) & (n == 32 ? 0xFFFFFFFFu : ((1u << n) - 1u));
        const uint32_t mine = ((sel >> lane) & 1u) ? zz : sym;
        const uint32_t src = __fns(keep, 0, lane + 1);
        const uint32_t moved = __shfl_sync(0xffffffffu, mine, src & 31u);
        n = __popc(keep);
        sym = lane < n ? moved : 0u;
    }
    if (lane < n) ids[lane] = (sym << 1) | (lane > 0);
    __syncwarp();
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
            uint32_t z; const uint32_t r = bpe_probe(t, nullptr, src[i], src[i + 1], z);
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

// kQueue: memo misses go to the warp's pending queue and are resolved 32 at a time, one word per lane (left-overs: warp per word);
// else they are resolved inside their tile (few: warp per word; many: one word per lane).
template <bool kNaive, bool kQueue>
struct BpeEncT {
    BpeTableDev t;
    using Stage = BpeStage;
    static constexpr bool kScratchLong = true;
    static constexpr bool kBatchSlowPath = kQueue;
    static constexpr bool kWarpShort = true;
    static constexpr bool kWarpLong = false;
    static constexpr bool kSplitCount = true;     // count pass = warm-up + leaf count kernel + resolve kernel (encode.cuh)
    __device__ __forceinline__ void stage_init(Stage &s) const {
        const uint4 *src = reinterpret_cast<const uint4 *>(t.stage);
        uint4 *dst = reinterpret_cast<uint4 *>(&s);
        for (uint32_t i = threadIdx.x; i < sizeof(Stage) / 16; i += blockDim.x) dst[i] = __ldg(src + i);
        __syncthreads();
    }
    // memo form of an id: the 16-bit symbol; the continuation bit is the position (token = symbol << 1 | k > 0, bpe.py:240-241),
    // so models of up to 65536 symbols are served by the 16-bit path
    __device__ static __forceinline__ bool narrow16(uint32_t id, uint32_t k, uint32_t &v16) { (void)k; v16 = id >> 1; return (id >> 1) < 65536u; }
    __device__ static __forceinline__ uint32_t expand16(uint32_t v16, uint32_t k) { return (v16 << 1) | (k > 0 ? 1u : 0u); }
    __device__ __forceinline__ uint32_t encode_short(const Stage *sg, const uint8_t *p, uint32_t nbytes, uint32_t *buf, uint32_t &h6) const {
        (void)h6;
        return bpe_encode_short<kNaive>(t, sg, p, nbytes, buf);
    }
    __device__ __forceinline__ uint32_t encode_short_warp(const Stage *sg, const uint8_t *p, uint32_t nbytes, uint32_t *ids) const {
        return bpe_encode_short_warp<kNaive>(t, sg, p, nbytes, ids);
    }
    __device__ __forceinline__ uint32_t encode_long_warp(const uint8_t *p, uint32_t nbytes, uint32_t *bufA, uint32_t *bufB,
                                                         uint32_t **result) const {
        return bpe_encode_long_warp<kNaive>(t, p, nbytes, bufA, bufB, result);
    }
};
using BpeEnc = BpeEncT<false, true>;          // FastBPE.encode_word
using BpeEncInTile = BpeEncT<false, false>;   // the same, misses resolved inside their tile (swt_tune("bpe_queue", 0))
using NaiveBpeEnc = BpeEncT<true, true>;      // NaiveBPE.encode_word (merge lists without repeated pairs)

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
    // shared-memory image: Bloom filter over every ranked pair + the lowest-rank merges, direct mapped (lowest rank keeps a slot)
    std::vector<BpeStage> stage(1);
    memset(stage.data(), 0, sizeof(BpeStage));
    for (uint32_t i = 0; i < kHotSlots; ++i) stage[0].hot[i] = make_uint4(0, 0, kEmptyRank, 0);
    {
        std::vector<uint4> ranked;
        for (const uint4 &e : slots) if (e.z != kEmptyRank) ranked.push_back(e);
        std::sort(ranked.begin(), ranked.end(), [](const uint4 &x, const uint4 &y) { return x.z < y.z; });
        for (const uint4 &e : ranked) {
            const uint64_t hh = mix64(((uint64_t)e.x << 32) | e.y);
            const uint32_t b1 = (uint32_t)(hh >> 32) & (kBloomWords * 32 - 1), b2 = (uint32_t)(hh >> 47) & (kBloomWords * 32 - 1);
            stage[0].bloom[b1 >> 5] |= 1u << (b1 & 31u); stage[0].bloom[b2 >> 5] |= 1u << (b2 & 31u);
            uint4 &hslot = stage[0].hot[(uint32_t)(hh >> 24) & (kHotSlots - 1)];
            if (hslot.z == kEmptyRank) hslot = e;
        }
    }
    std::vector<uint32_t> lut(65536, 0xFFFFFFFFu), hi_cp, hi_id;
    for (uint32_t i = 0; i < n_chars; ++i) {
        SWT_REQUIRE(i == 0 || h_char_cp[i] > h_char_cp[i - 1], "char_cp must be strictly ascending");
        if (h_char_cp[i] < 0x10000u) lut[h_char_cp[i]] = h_char_id[i];
        else { hi_cp.push_back(h_char_cp[i]); hi_id.push_back(h_char_id[i]); }
    }
    // one device blob: slots | lut | hi_cp | hi_id | m_left | m_right | m_new
    Carver sz(nullptr);
    sz.take<BpeStage>(1); sz.take<uint4>(n_slots); sz.take<uint32_t>(65536); sz.take<uint32_t>(hi_cp.size() + 1); sz.take<uint32_t>(hi_id.size() + 1);
    sz.take<uint32_t>(n_merges + 1); sz.take<uint32_t>(n_merges + 1); sz.take<uint32_t>(n_merges + 1);
    void *blob = nullptr;
    SWT_CUDA_OK(cudaMalloc(&blob, sz.used()));
    Carver cv(blob);
    BpeStage *d_stage = cv.take<BpeStage>(1);
    uint4 *d_slots = cv.take<uint4>(n_slots);
    uint32_t *d_lut = cv.take<uint32_t>(65536);
    uint32_t *d_hicp = cv.take<uint32_t>(hi_cp.size() + 1), *d_hiid = cv.take<uint32_t>(hi_id.size() + 1);
    uint32_t *d_l = cv.take<uint32_t>(n_merges + 1), *d_r = cv.take<uint32_t>(n_merges + 1), *d_n = cv.take<uint32_t>(n_merges + 1);
    cudaError_t e = cudaMemcpy(d_slots, slots.data(), n_slots * sizeof(uint4), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_stage, stage.data(), sizeof(BpeStage), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_lut, lut.data(), 65536 * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !hi_cp.empty()) e = cudaMemcpy(d_hicp, hi_cp.data(), hi_cp.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !hi_id.empty()) e = cudaMemcpy(d_hiid, hi_id.data(), hi_id.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n_merges) e = cudaMemcpy(d_l, h_left, n_merges * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n_merges) e = cudaMemcpy(d_r, h_right, n_merges * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n_merges) e = cudaMemcpy(d_n, h_merged, n_merges * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(blob); set_error(std::string("table upload: ") + cudaGetErrorString(e)); return SWT_ERR_CUDA; }
    swt_bpe_table *t = new swt_bpe_table();
    t->dev = BpeTableDev{d_stage, d_slots, (uint32_t)(n_slots - 1), d_lut, d_hicp, d_hiid, (uint32_t)hi_cp.size(), d_l, d_r, d_n, n_merges};
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
    if (!g_tune.bpe_queue)
        return launch_encode_tiles(BpeEncInTile{t->dev}, d_arena, d_word_off, n_words, long_word_bytes, d_out_ids, out_cap, d_out_tok_off,
                                   tok_base, d_workspace, workspace_bytes, d_status, st);
    return launch_encode_tiles(BpeEnc{t->dev}, d_arena, d_word_off, n_words, long_word_bytes, d_out_ids, out_cap, d_out_tok_off,
                               tok_base, d_workspace, workspace_bytes, d_status, st);
}
}  // namespace swt

namespace swt {
int bpe_small_launch(const swt_bpe_table *t, int naive, const pt::PretokDev &pd, bool bert, const uint8_t *h_text, const uint8_t *d_text, uint32_t n, const SmallArgs &a,
                     cudaStream_t st) {
    SWT_REQUIRE(t != nullptr, "NULL table");
    if (naive) return launch_tokenize_small(NaiveBpeEnc{t->dev}, pd, bert, h_text, d_text, n, a, st);
    return launch_tokenize_small(BpeEnc{t->dev}, pd, bert, h_text, d_text, n, a, st);
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
