// wp_encode.cu -- HP-2: FastWP.tokenize / matchloop (reference source/wordpiece.py:233-316) over a
// flattened WPTrie_E2E (reference source/utils.py:66-139) on sm_100a.
//
// Host side (swt_wp_trie_create): builds the trie in index arrays, runs the failure-link / failure-pop
// precompute of utils.py:108-139 breadth first from [root, root_sharp], and flattens it into
//   edge table   open-addressing hash, 16-byte slots {node, code point, child, -}  (one 128-bit load/probe)
//   root_lut     direct child table of the root for code points < kRootLut
//   node_info    16 bytes per node {failure link, pops offset, pops count, vocabulary id of the entry ending here}
//   pops         token ids emitted on a failure transition
//   alnum        Python str.isalnum bitmap (wordpiece.py:287-288, utils.py:137)
// all small enough to stay L2/L1 resident (about 3 MB for a 20K vocabulary).
//
// Device side (tile kernel and word-type memo: encode.cuh): one thread walks one whitespace-free chunk (goto +
// failure transitions); chunks longer than kShortBytes are walked twice (count, then write).
#include <algorithm>
#include <deque>
#include <unordered_map>
#include <vector>

#include "encode.cuh"

namespace swt {

constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr uint32_t kRootLut = 2048;
constexpr uint32_t kNodeRoot = 0, kNodeRootP = 1, kNodeRootSharp = 3;   // insert("##") creates nodes 2 and 3

struct WpTrieDev {
    const uint4 *edges; uint32_t edge_mask;
    const uint32_t *root_lut;
    const uint4 *node_info;
    const uint32_t *pops;
    const uint32_t *alnum;       // 0x110000 bits
    uint32_t n_vocab;
    uint32_t sharp_special[2]; uint32_t n_sharp_special;
};

}  // namespace swt

struct swt_wp_trie {
    swt::WpTrieDev dev;
    int device;
    void *d_blob;
    uint64_t n_nodes, n_edges, n_pops, n_rootp;
};

namespace swt {

__device__ __forceinline__ uint32_t wp_edge(const WpTrieDev &t, uint32_t node, uint32_t cp) {
    if (node == kNodeRoot && cp < kRootLut) return __ldg(&t.root_lut[cp]);
    uint32_t h = (uint32_t)mix64(((uint64_t)node << 32) | cp) & t.edge_mask;
    for (;;) {
        uint4 e = __ldg(&t.edges[h]);
        if (e.x == kNone) return kNone;
        if (e.x == node && e.y == cp) return e.z;
        h = (h + 1) & t.edge_mask;
    }
}
__device__ __forceinline__ bool wp_alnum(const WpTrieDev &t, uint32_t cp) {
    // ASCII without a table access (str.isalnum of an ASCII character: 0-9, A-Z, a-z): the bitmap load would sit on the dependent chain
    // of every character of the trie walk
    if (cp < 0x80u) return (cp - 0x30u) < 10u || ((cp | 0x20u) - 0x61u) < 26u;
    return cp < 0x110000u && ((__ldg(&t.alnum[cp >> 5]) >> (cp & 31u)) & 1u);
}

struct CountEmit {                       // counting pass over a long chunk
    uint32_t n = 0;
    __device__ __forceinline__ void push(uint32_t) { ++n; }
    __device__ __forceinline__ uint32_t size() const { return n; }
    __device__ __forceinline__ void truncate(uint32_t k) { n = k; }
};
struct ArrayEmit {                       // ids into a buffer of `cap` entries (thread-local, shared or global)
    uint32_t *out; uint32_t cap; uint32_t n = 0;
    __device__ __forceinline__ void push(uint32_t tok) { if (n < cap) out[n] = tok; ++n; }
    __device__ __forceinline__ uint32_t size() const { return n; }
    __device__ __forceinline__ void truncate(uint32_t k) { n = k; }
};

// One iteration of the loop of FastWP.tokenize (wordpiece.py:251-269) on s = chunk + " ": match loop from position i (a word
// boundary), accept / reject, advance to the next boundary.  prev_punct = ispunc(s[i-1]) (false at i == 0) on entry and on exit.
// Returns the new position.  A segment is a function of its start position alone, which is what lets the lanes of a warp walk the
// segments of a long chunk independently (wp_long_count_warp).
template <class Emit>
__device__ __forceinline__ uint32_t wp_encode_segment(const WpTrieDev &t, const uint8_t *p, uint32_t nbytes, uint32_t i, bool &prev_punct,
                                                      Emit &emit, uint32_t &h6) {
    const uint32_t seg = emit.size(), i0 = i;
    uint32_t node = kNodeRoot, cp = 0, adv = 1;
    // ---- matchloop (wordpiece.py:291-316)
    for (;;) {
        if (i < nbytes) cp = utf8_decode(p + i, nbytes - i, adv); else { cp = 0x20u; adv = 1; }   // the appended " "
        uint32_t child = (i < nbytes) ? wp_edge(t, node, cp) : kNone;   // no vocabulary entry holds whitespace
        bool returned = false;
        while (child == kNone) {                                        // failure transitions :308-312
            const uint4 info = __ldg(&t.node_info[node]);
            if (info.x == kNone) { returned = true; break; }
            for (uint32_t k = 0; k < info.z; ++k) emit.push(__ldg(&t.pops[info.y + k]));
            node = info.x;
            child = (i < nbytes) ? wp_edge(t, node, cp) : kNone;
        }
        if (returned) break;
        node = child; i += adv; prev_punct = !wp_alnum(t, cp);          // goto transition :314-315
    }
    // ---- accept / reject (wordpiece.py:255-261); cp is s[i]
    bool bnd = prev_punct || i >= nbytes || !wp_alnum(t, cp);           // iswdbndry :285
    const bool ok = bnd && (node == kNodeRoot || node == kNodeRootSharp || node == kNodeRootP);
    if (!ok) { emit.truncate(seg); emit.push(t.n_vocab); }              // "['UNK']"
    else if (node == kNodeRootSharp && emit.size() == seg) {
        for (uint32_t k = 0; k < t.n_sharp_special; ++k) emit.push(t.sharp_special[k]);
    }
    // ---- advance to the next boundary (:265-266), then skip whitespace (:268-269)
    while (!bnd) {
        i += adv; prev_punct = false;                                   // s[i] was alnum here
        if (i < nbytes) { cp = utf8_decode(p + i, nbytes - i, adv); bnd = !wp_alnum(t, cp); }
        else bnd = true;
    }
    if (i == i0) {
        // H6: punctuation that is not a child of the root -- the reference spins forever here.
        // Documented extension: emit "['UNK']" and advance one character.
        ++h6; emit.push(t.n_vocab);
        i += adv; prev_punct = !wp_alnum(t, cp);
    }
    return i;
}

// The segments from a word boundary i up to the NEXT word boundary.  Normally that is one segment; only the H6 extension can leave
// the position inside a run of letters (it advances one character after a boundary at which nothing matched), and the segments
// that follow until the next boundary belong to the same unit of work.
template <class Emit>
__device__ __forceinline__ uint32_t wp_encode_to_boundary(const WpTrieDev &t, const uint8_t *p, uint32_t nbytes, uint32_t i, bool &prev_punct,
                                                          Emit &emit, uint32_t &h6) {
    for (;;) {
        i = wp_encode_segment(t, p, nbytes, i, prev_punct, emit, h6);
        if (i >= nbytes || prev_punct) return i;
        uint32_t adv;
        if (!wp_alnum(t, utf8_decode(p + i, nbytes - i, adv))) return i;
    }
}

// FastWP.tokenize on s = chunk + " " (wordpiece.py:248-269); the chunk holds no whitespace.
template <class Emit>
__device__ __forceinline__ void wp_encode_chunk(const WpTrieDev &t, const uint8_t *p, uint32_t nbytes, Emit &emit, uint32_t &h6) {
    uint32_t i = 0;
    bool prev_punct = false;                       // ispunc(s[i-1]); false at i == 0
    while (i < nbytes) i = wp_encode_segment(t, p, nbytes, i, prev_punct, emit, h6);   // the virtual space ends the chunk
}

// ---- long chunks, split across the lanes of a warp (north-star item 3) ----------------------------------------------------------
// Every iteration of the loop above starts at a word boundary (iswdbndry: the previous or the current character is punctuation) with
// the trie at its root, so the segment starting at a boundary does not depend on what came before.  The warp lists the boundary
// positions of the chunk, the lanes walk the segment of every boundary (count only; "segment" = up to the next boundary, see
// wp_encode_to_boundary), and the segments are chained from position 0:
// with a vocabulary whose entries do not mix punctuation and letters every segment ends at the next boundary and all of them are
// used (checked in parallel); otherwise (an entry like "a.b" lets a match loop run across a boundary) lane 0 follows the chain and
// the segments it skips are dropped, which is exactly what the sequential loop does.  A run of letters without punctuation is one
// segment and stays sequential (the trie state carries).
// Scratch (u32): [0] tokens of the chunk, [1] number of boundaries; record c at 16 + 4c: position, end -> output offset (or kNone
// when the chain skips it), tokens, H6 events.
constexpr uint32_t kSegRec = 4;
__device__ __forceinline__ bool wp_prev_is_punct(const WpTrieDev &t, const uint8_t *p, uint32_t nbytes, uint32_t i) {
    if (i == 0) return false;
    uint32_t j = i - 1;
    while (j > 0 && (p[j] & 0xC0u) == 0x80u) --j;
    uint32_t adv;
    return !wp_alnum(t, utf8_decode(p + j, nbytes - j, adv));
}
static __device__ __noinline__ uint32_t wp_long_count_warp(const WpTrieDev &t, const uint8_t *p, uint32_t nbytes, uint32_t *scratch, uint32_t &h6_out) {
    const uint32_t lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
    uint32_t *rec = scratch + kLongHeader;
    // a. boundary positions, in ascending order
    uint32_t nc = 0;
    bool carry_alnum = true;                                              // class of the last character of the previous 32 bytes
    for (uint32_t base = 0; base < nbytes; base += 32) {
        const uint32_t b = base + lane;
        const bool start = b < nbytes && (p[b] & 0xC0u) != 0x80u;
        bool aln = false;
        if (start) { uint32_t adv; aln = wp_alnum(t, utf8_decode(p + b, nbytes - b, adv)); }
        const uint32_t S = __ballot_sync(0xffffffffu, start), A = __ballot_sync(0xffffffffu, aln);
        const uint32_t below = S & lt;
        const bool prev_alnum = below ? ((A >> (31 - __clz(below))) & 1u) : carry_alnum;
        const bool cand = start && (b == 0 || !prev_alnum || !aln);
        const uint32_t cm = __ballot_sync(0xffffffffu, cand);
        if (cand) rec[kSegRec * (nc + __popc(cm & lt))] = b;
        nc += __popc(cm);
        if (S) carry_alnum = (A >> (31 - __clz(S))) & 1u;
    }
    __syncwarp();
    // b. one segment per boundary, counting only
    for (uint32_t c = lane; c < nc; c += 32) {
        const uint32_t i = rec[kSegRec * c];
        bool prev_punct = wp_prev_is_punct(t, p, nbytes, i);
        CountEmit e; uint32_t sh6 = 0;
        const uint32_t end = wp_encode_to_boundary(t, p, nbytes, i, prev_punct, e, sh6);
        rec[kSegRec * c + 1] = end; rec[kSegRec * c + 2] = e.n; rec[kSegRec * c + 3] = sh6;
    }
    __syncwarp();
    // c. chain: every segment ends where the next boundary is?
    bool ok = true;
    for (uint32_t c = lane; c < nc; c += 32) {
        const uint32_t end = rec[kSegRec * c + 1];
        ok = ok && (c + 1 < nc ? end == rec[kSegRec * (c + 1)] : end >= nbytes);
    }
    ok = __all_sync(0xffffffffu, ok);
    uint32_t total = 0, h6 = 0;
    if (ok) {                                                             // all used: offsets = exclusive scan of the token counts
        for (uint32_t c0 = 0; c0 < nc; c0 += 32) {
            const uint32_t c = c0 + lane;
            const uint32_t n = c < nc ? rec[kSegRec * c + 2] : 0u;
            h6 += c < nc ? rec[kSegRec * c + 3] : 0u;
            uint32_t incl = n;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += v; }
            if (c < nc) rec[kSegRec * c + 1] = total + incl - n;
            total += __shfl_sync(0xffffffffu, incl, 31);
        }
    } else {
        if (lane == 0) {
            uint32_t pos = 0, c = 0;
            while (pos < nbytes && c < nc) {
                while (c < nc && rec[kSegRec * c] < pos) { rec[kSegRec * c + 1] = kNone; ++c; }      // skipped by the chain
                if (c >= nc || rec[kSegRec * c] != pos) break;                                       // cannot happen: an end is a boundary
                pos = rec[kSegRec * c + 1];
                rec[kSegRec * c + 1] = total; total += rec[kSegRec * c + 2]; h6 += rec[kSegRec * c + 3];
                ++c;
            }
            for (; c < nc; ++c) rec[kSegRec * c + 1] = kNone;
        }
        total = __shfl_sync(0xffffffffu, total, 0);
    }
    if (lane == 0) { scratch[0] = total; scratch[1] = nc; }
    __syncwarp();
    h6_out = h6;                                                          // the lanes' shares add up to the chunk's H6 events
    return total;
}
static __device__ __noinline__ void wp_long_emit_warp(const WpTrieDev &t, const uint8_t *p, uint32_t nbytes, const uint32_t *scratch, uint32_t *dst,
                                                      uint32_t n_total) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t *rec = scratch + kLongHeader;
    const uint32_t nc = scratch[1];
    for (uint32_t c = lane; c < nc; c += 32) {
        const uint32_t off = rec[kSegRec * c + 1];
        if (off == kNone || off > n_total) continue;
        const uint32_t i = rec[kSegRec * c];
        bool prev_punct = wp_prev_is_punct(t, p, nbytes, i);
        ArrayEmit e{dst + off, min(rec[kSegRec * c + 2], n_total - off)};
        uint32_t dummy = 0;
        wp_encode_to_boundary(t, p, nbytes, i, prev_punct, e, dummy);
    }
    __syncwarp();
}

// NaiveWP.encode_word (wordpiece.py:131-158): greedy longest prefix in the vocabulary; the rest is looked up as "##" + rest.
// The whole word becomes "[UNK]" (id n_vocab + 1) when a piece has no match.  A continuation piece must match at least one
// character beyond its "##" (a bare "#"/"##" match makes the reference's remainder grow forever; DESIGN.md, parity domain).
template <class Emit>
__device__ __forceinline__ void wp_naive_encode_word(const WpTrieDev &t, const uint8_t *p, uint32_t nbytes, Emit &emit) {
    const uint32_t seg = emit.size();
    uint32_t i = 0;
    bool first = true;
    while (i < nbytes) {
        uint32_t node = first ? kNodeRoot : kNodeRootSharp, best = kNone, best_end = i, j = i;
        while (j < nbytes) {
            uint32_t adv;
            const uint32_t cp = utf8_decode(p + j, nbytes - j, adv);
            const uint32_t child = wp_edge(t, node, cp);
            if (child == kNone) break;
            node = child; j += adv;
            const uint32_t tok = __ldg(&t.node_info[node]).w;                // vocabulary id of the entry ending here, or kNone
            if (tok != kNone) { best = tok; best_end = j; }
        }
        if (best == kNone) { emit.truncate(seg); emit.push(t.n_vocab + 1); return; }
        emit.push(best); i = best_end; first = false;
    }
}

struct WpStage { uint32_t unused; };      // the trie stays in L1/L2: nothing is staged in shared memory

struct NaiveWpEnc {
    WpTrieDev t;
    using Stage = WpStage;
    static constexpr bool kScratchLong = false;
    static constexpr bool kBatchSlowPath = true;
    static constexpr bool kWarpShort = false;
    static constexpr bool kWarpLong = false;
    static constexpr bool kSplitCount = false;
    __device__ __forceinline__ void stage_init(Stage &) const {}
    __device__ static __forceinline__ bool narrow16(uint32_t id, uint32_t k, uint32_t &v16) { (void)k; v16 = id; return id < 65536u; }
    __device__ static __forceinline__ uint32_t expand16(uint32_t v16, uint32_t k) { (void)k; return v16; }
    __device__ __forceinline__ uint32_t encode_short(const Stage *, const uint8_t *p, uint32_t nbytes, uint32_t *buf, uint32_t &h6) const {
        (void)h6;
        ArrayEmit e{buf, (uint32_t)kShortBytes};
        wp_naive_encode_word(t, p, nbytes, e);
        return e.n;
    }
    __device__ __forceinline__ uint32_t long_count(const uint8_t *p, uint32_t nbytes, uint32_t &h6) const {
        (void)h6;
        CountEmit e;
        wp_naive_encode_word(t, p, nbytes, e);
        return e.n;
    }
    __device__ __forceinline__ void long_emit(const uint8_t *p, uint32_t nbytes, uint32_t *dst, uint32_t cap, uint32_t &h6) const {
        (void)h6;
        ArrayEmit e{dst, cap};
        wp_naive_encode_word(t, p, nbytes, e);
    }
};

struct WpEnc {
    WpTrieDev t;
    using Stage = WpStage;
    static constexpr bool kScratchLong = false;
    static constexpr bool kBatchSlowPath = true;
    static constexpr bool kWarpShort = false;
    static constexpr bool kWarpLong = true;
    static constexpr bool kSplitCount = false;    // the single count kernel is faster for this encoder (encode.cuh)
    __device__ __forceinline__ void stage_init(Stage &) const {}
    __device__ static __forceinline__ bool narrow16(uint32_t id, uint32_t k, uint32_t &v16) { (void)k; v16 = id; return id < 65536u; }
    __device__ static __forceinline__ uint32_t expand16(uint32_t v16, uint32_t k) { (void)k; return v16; }
    // long chunks with scratch: the segments of the chunk are walked by the lanes of the warp (all 32 lanes call these)
    __host__ __device__ static __forceinline__ unsigned long long long_scratch_need(uint32_t nbytes) {
        return (kLongHeader + (unsigned long long)kSegRec * nbytes + 15ull) & ~15ull;
    }
    __device__ __forceinline__ uint32_t long_count_warp(const uint8_t *p, uint32_t nbytes, uint32_t *scratch, uint32_t &h6) const {
        return wp_long_count_warp(t, p, nbytes, scratch, h6);
    }
    __device__ __forceinline__ void long_emit_warp(const uint8_t *p, uint32_t nbytes, const uint32_t *scratch, uint32_t *dst, uint32_t n) const {
        wp_long_emit_warp(t, p, nbytes, scratch, dst, n);
    }
    __device__ __forceinline__ uint32_t encode_short(const Stage *, const uint8_t *p, uint32_t nbytes, uint32_t *buf, uint32_t &h6) const {
        ArrayEmit e{buf, (uint32_t)kShortBytes};
        wp_encode_chunk(t, p, nbytes, e, h6);
        return e.n;
    }
    // chunks longer than kShortBytes are walked twice: count, then write at the final position
    __device__ __forceinline__ uint32_t long_count(const uint8_t *p, uint32_t nbytes, uint32_t &h6) const {
        CountEmit e;
        wp_encode_chunk(t, p, nbytes, e, h6);
        return e.n;
    }
    __device__ __forceinline__ void long_emit(const uint8_t *p, uint32_t nbytes, uint32_t *dst, uint32_t cap, uint32_t &h6) const {
        ArrayEmit e{dst, cap};
        wp_encode_chunk(t, p, nbytes, e, h6);
    }
};

// ---- host: trie construction + precompute ---------------------------------------------------------------
struct HostTrie {
    std::vector<uint32_t> ch, fail, token;         // per node: edge character, failure link, vocab id (or kNone)
    std::vector<uint8_t> is_end;
    std::vector<std::vector<uint32_t>> pops;
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> kids;   // (code point, child) in insertion order
    std::unordered_map<uint64_t, uint32_t> edge;                    // node<<21 | cp -> child

    uint32_t add_node(uint32_t c) {
        ch.push_back(c); fail.push_back(kNone); token.push_back(kNone); is_end.push_back(0);
        pops.emplace_back(); kids.emplace_back();
        return (uint32_t)ch.size() - 1;
    }
    uint32_t child(uint32_t node, uint32_t cp) const {
        auto it = edge.find(((uint64_t)node << 21) | cp);
        return it == edge.end() ? kNone : it->second;
    }
    uint32_t insert(const uint32_t *w, uint32_t len, uint32_t tok) {          // utils.py:87-105
        uint32_t node = kNodeRoot;
        for (uint32_t i = 0; i < len; ++i) {
            uint32_t c = child(node, w[i]);
            if (c == kNone) {
                c = add_node(w[i]);
                edge.emplace(((uint64_t)node << 21) | w[i], c);
                kids[node].emplace_back(w[i], c);
            }
            node = c;
        }
        is_end[node] = 1;
        if (tok != kNone) token[node] = tok;
        return node;
    }
};

}  // namespace swt

using namespace swt;

SWT_API int swt_wp_trie_create(const uint32_t *h_vocab_cps, const uint64_t *h_vocab_off, uint32_t n_vocab,
                               const uint8_t *h_alnum_bitmap, const uint32_t *h_sharp_special, uint32_t n_sharp_special,
                               int device, swt_wp_trie **out) {
    SWT_REQUIRE(out && h_alnum_bitmap && h_vocab_off, "NULL argument");
    SWT_REQUIRE(n_sharp_special <= 2, "NaiveWP.encode_word('##') longer than 2 tokens is outside the supported domain");
    SWT_REQUIRE(n_vocab < 0x7FFFFFF0u, "vocabulary too large");
    SWT_CUDA_OK(cudaSetDevice(device));
    auto alnum = [&](uint32_t cp) { return cp < 0x110000u && ((h_alnum_bitmap[cp >> 3] >> (cp & 7)) & 1); };

    HostTrie T;
    T.add_node(0);                                   // root
    T.add_node(0);                                   // root_p: no children, no failure link (utils.py:79)
    const uint32_t sharp[2] = {'#', '#'};
    const uint32_t r_sharp = T.insert(sharp, 2, kNone);                       // utils.py:81
    if (r_sharp != kNodeRootSharp) { set_error("internal: root_sharp id"); return SWT_ERR_INTERNAL; }
    for (uint32_t v = 0; v < n_vocab; ++v) {
        const uint32_t len = (uint32_t)(h_vocab_off[v + 1] - h_vocab_off[v]);
        for (uint32_t k = 0; k < len; ++k) SWT_REQUIRE(h_vocab_cps[h_vocab_off[v] + k] < 0x110000u, "code point out of range");
        T.insert(h_vocab_cps + h_vocab_off[v], len, v);                       // utils.py:83-84
    }
    // failure links and pops, breadth first from [root, root_sharp] (utils.py:113-139)
    std::deque<uint32_t> queue{kNodeRoot, kNodeRootSharp};
    while (!queue.empty()) {
        const uint32_t cur = queue.front(); queue.pop_front();
        for (const auto &kc : T.kids[cur]) {
            const uint32_t c = kc.first, child = kc.second;
            if (child == kNodeRootSharp) continue;
            if (T.is_end[child]) {                   // a vocabulary entry ends here: pop it, continue from "##"
                T.fail[child] = kNodeRootSharp;
                T.pops[child].assign(1, T.token[child]);
            } else {
                uint32_t f = T.fail[cur];
                std::vector<uint32_t> extra;
                while (f != kNone && T.child(f, c) == kNone) {
                    extra.insert(extra.end(), T.pops[f].begin(), T.pops[f].end());
                    f = T.fail[f];
                }
                if (f != kNone) {
                    T.fail[child] = T.child(f, c);
                    T.pops[child] = T.pops[cur];
                    T.pops[child].insert(T.pops[child].end(), extra.begin(), extra.end());
                }
            }
            if (!alnum(T.ch[child])) T.fail[child] = kNodeRootP;              // utils.py:137-138
            queue.push_back(child);
        }
    }
    // flatten
    const uint64_t n_nodes = T.ch.size(), n_edges = T.edge.size();
    const uint64_t n_slots = next_pow2(n_edges * 2 + 16);
    std::vector<uint4> slots(n_slots, make_uint4(kNone, 0, 0, 0));
    std::vector<uint32_t> root_lut(kRootLut, kNone);
    for (uint32_t node = 0; node < n_nodes; ++node)
        for (const auto &kc : T.kids[node]) {
            uint64_t h = mix64(((uint64_t)node << 32) | kc.first) & (n_slots - 1);
            while (slots[h].x != kNone) h = (h + 1) & (n_slots - 1);
            slots[h] = make_uint4(node, kc.first, kc.second, 0);
            if (node == kNodeRoot && kc.first < kRootLut) root_lut[kc.first] = kc.second;
        }
    std::vector<uint4> info(n_nodes);
    std::vector<uint32_t> pops;
    uint64_t n_rootp = 0;
    for (uint32_t node = 0; node < n_nodes; ++node) {
        info[node] = make_uint4(T.fail[node], (uint32_t)pops.size(), (uint32_t)T.pops[node].size(), T.token[node]);
        pops.insert(pops.end(), T.pops[node].begin(), T.pops[node].end());
        n_rootp += (T.fail[node] == kNodeRootP);
    }
    Carver sz(nullptr);
    sz.take<uint4>(n_slots); sz.take<uint4>(n_nodes); sz.take<uint32_t>(kRootLut); sz.take<uint32_t>(pops.size() + 1); sz.take<uint32_t>(0x110000 / 32);
    void *blob = nullptr;
    SWT_CUDA_OK(cudaMalloc(&blob, sz.used()));
    Carver cv(blob);
    uint4 *d_slots = cv.take<uint4>(n_slots); uint4 *d_info = cv.take<uint4>(n_nodes);
    uint32_t *d_lut = cv.take<uint32_t>(kRootLut); uint32_t *d_pops = cv.take<uint32_t>(pops.size() + 1);
    uint32_t *d_alnum = cv.take<uint32_t>(0x110000 / 32);
    cudaError_t e = cudaMemcpy(d_slots, slots.data(), n_slots * sizeof(uint4), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_info, info.data(), n_nodes * sizeof(uint4), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_lut, root_lut.data(), kRootLut * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !pops.empty()) e = cudaMemcpy(d_pops, pops.data(), pops.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_alnum, h_alnum_bitmap, 0x110000 / 8, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(blob); set_error(std::string("trie upload: ") + cudaGetErrorString(e)); return SWT_ERR_CUDA; }
    swt_wp_trie *t = new swt_wp_trie();
    t->dev.edges = d_slots; t->dev.edge_mask = (uint32_t)(n_slots - 1); t->dev.root_lut = d_lut; t->dev.node_info = d_info;
    t->dev.pops = d_pops; t->dev.alnum = d_alnum; t->dev.n_vocab = n_vocab; t->dev.n_sharp_special = n_sharp_special;
    for (uint32_t k = 0; k < 2; ++k) t->dev.sharp_special[k] = k < n_sharp_special ? h_sharp_special[k] : 0;
    t->device = device; t->d_blob = blob;
    t->n_nodes = n_nodes; t->n_edges = n_edges; t->n_pops = pops.size(); t->n_rootp = n_rootp;
    *out = t;
    return SWT_OK;
}

SWT_API void swt_wp_trie_destroy(swt_wp_trie *t) {
    if (!t) return;
    cudaSetDevice(t->device);
    cudaFree(t->d_blob);
    delete t;
}

SWT_API int swt_wp_trie_stats(const swt_wp_trie *t, uint64_t *n_nodes, uint64_t *n_edges, uint64_t *n_pops, uint64_t *n_rootp) {
    SWT_REQUIRE(t != nullptr, "NULL trie");
    if (n_nodes) *n_nodes = t->n_nodes;
    if (n_edges) *n_edges = t->n_edges;
    if (n_pops) *n_pops = t->n_pops;
    if (n_rootp) *n_rootp = t->n_rootp;
    return SWT_OK;
}

namespace swt {
int wp_encode_launch(const swt_wp_trie *t, const uint8_t *d_arena, const uint32_t *d_word_off, uint32_t n_words, uint64_t long_word_bytes,
                     uint32_t *d_out_ids, uint64_t out_cap, uint32_t *d_out_tok_off, uint32_t tok_base,
                     void *d_workspace, size_t workspace_bytes, uint32_t *d_status, cudaStream_t st) {
    SWT_REQUIRE(t != nullptr, "NULL trie");
    return launch_encode_tiles(WpEnc{t->dev}, d_arena, d_word_off, n_words, long_word_bytes, d_out_ids, out_cap, d_out_tok_off, tok_base,
                               d_workspace, workspace_bytes, d_status, st);
}
}  // namespace swt

namespace swt {
int wp_small_launch(const swt_wp_trie *t, int naive, const pt::PretokDev &pd, bool bert, const uint8_t *h_text, const uint8_t *d_text, uint32_t n, const SmallArgs &a,
                     cudaStream_t st) {
    SWT_REQUIRE(t != nullptr, "NULL trie");
    if (naive) return launch_tokenize_small(NaiveWpEnc{t->dev}, pd, bert, h_text, d_text, n, a, st);
    return launch_tokenize_small(WpEnc{t->dev}, pd, bert, h_text, d_text, n, a, st);
}
}  // namespace swt

SWT_API int swt_wp_encode(const swt_wp_trie *t, const uint8_t *d_arena, const uint32_t *d_word_off, uint32_t n_words,
                          uint64_t long_word_bytes, uint32_t *d_out_ids, uint64_t out_cap, uint32_t *d_out_tok_off,
                          void *d_workspace, size_t workspace_bytes, uint32_t *d_status, void *stream) {
    // long_word_bytes > 0: chunks longer than 32 bytes are split into their segments across the lanes of a warp (scratch for the
    // segment records); 0: the owning lane walks such a chunk alone
    return wp_encode_launch(t, d_arena, d_word_off, n_words, long_word_bytes, d_out_ids, out_cap, d_out_tok_off, 0u, d_workspace,
                            workspace_bytes, d_status, (cudaStream_t)stream);
}

SWT_API int swt_wp_encode_naive(const swt_wp_trie *t, const uint8_t *d_arena, const uint32_t *d_word_off, uint32_t n_words,
                                uint64_t long_word_bytes, uint32_t *d_out_ids, uint64_t out_cap, uint32_t *d_out_tok_off,
                                void *d_workspace, size_t workspace_bytes, uint32_t *d_status, void *stream) {
    (void)long_word_bytes;
    SWT_REQUIRE(t != nullptr, "NULL trie");
    return launch_encode_tiles(NaiveWpEnc{t->dev}, d_arena, d_word_off, n_words, 0, d_out_ids, out_cap, d_out_tok_off, 0u, d_workspace,
                               workspace_bytes, d_status, (cudaStream_t)stream);
}
