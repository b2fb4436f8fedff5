// pretok.cuh -- device side of the pre-tokenizer (see pretok.cu): per-word analysis and the count / write passes over one 4 KiB tile.
// Shared by the grid kernels of pretok.cu and the single-CTA small-call kernel (encode.cuh, tokenize_small_kernel).
#pragma once
#include <algorithm>

#include "common.cuh"

struct swt_pretok {
    int device; int mode;                // SWT_PRETOK_PYTHON_SPLIT / SWT_PRETOK_BERT
    uint32_t *d_lower; uint32_t n_lower;
    uint32_t *d_multi; uint32_t n_multi;
    uint8_t *d_cased, *d_ignorable;      // 0x110000 / 8 bytes each, or nullptr (text must then not contain U+03A3)
};

namespace swt {
namespace pt {

constexpr uint32_t kTileBytes = 4096, kGroupTiles = 1024;
constexpr uint32_t kSmallTiles = 8;      // small-call path: texts of up to kSmallTiles tiles are pre-tokenized by one CTA, one warp per tile
constexpr uint32_t kLowerMulti = 0x80000000u, kLowerSigma = 0x40000000u, kLowerPunct = 0x20000000u, kLowerCpMask = 0x1FFFFFu;
enum { kPtCode = 0, kPtWords = 1, kPtBytesLo = 2, kPtBytesHi = 3 };

struct PretokDev {
    const uint32_t *lower; uint32_t n_lower;
    const uint32_t *multi;
    const uint8_t *cased, *ignorable;
};
struct PretokWs { unsigned long long *tile_sum; unsigned long long *group_base; uint32_t n_tiles, n_groups; };

inline size_t pretok_layout(uint64_t n_bytes, void *base, PretokWs *ws) {
    Carver c(base);
    const uint64_t n_tiles = (n_bytes + kTileBytes - 1) / kTileBytes, n_groups = (n_tiles + kGroupTiles - 1) / kGroupTiles;
    ws->tile_sum = c.take<unsigned long long>(n_tiles + 1);
    ws->group_base = c.take<unsigned long long>(n_groups + 1);
    ws->n_tiles = (uint32_t)n_tiles; ws->n_groups = (uint32_t)n_groups;
    return c.used();
}

__device__ __forceinline__ bool is_space_ascii(uint32_t b) { return b == 0x20u || (b - 0x09u) < 5u || (b - 0x1Cu) < 4u; }
// the non-ASCII str.isspace() characters: U+0085 U+00A0 | U+1680 | U+2000-200A U+2028 U+2029 U+202F | U+205F | U+3000
__device__ __forceinline__ bool is_space_2(uint32_t b0, uint32_t b1) { return b0 == 0xC2u && (b1 == 0x85u || b1 == 0xA0u); }
__device__ __forceinline__ bool is_space_3(uint32_t b0, uint32_t b1, uint32_t b2) {
    if (b0 == 0xE2u) return (b1 == 0x80u && ((b2 - 0x80u) <= 0x0Au || b2 == 0xA8u || b2 == 0xA9u || b2 == 0xAFu)) || (b1 == 0x81u && b2 == 0x9Fu);
    return (b0 == 0xE1u && b1 == 0x9Au && b2 == 0x80u) || (b0 == 0xE3u && b1 == 0x80u && b2 == 0x80u);
}
__device__ __forceinline__ uint32_t utf8_len_of(uint32_t cp) { return cp < 0x80u ? 1u : cp < 0x800u ? 2u : cp < 0x10000u ? 3u : 4u; }
__device__ __forceinline__ bool bitmap_bit(const uint8_t *bm, uint32_t cp) { return cp < 0x110000u && ((bm[cp >> 3] >> (cp & 7)) & 1u); }
__device__ __forceinline__ uint8_t *put_utf8(uint8_t *d, uint32_t cp) {
    if (cp < 0x80u) { *d++ = (uint8_t)cp; }
    else if (cp < 0x800u) { *d++ = (uint8_t)(0xC0u | (cp >> 6)); *d++ = (uint8_t)(0x80u | (cp & 0x3Fu)); }
    else if (cp < 0x10000u) { *d++ = (uint8_t)(0xE0u | (cp >> 12)); *d++ = (uint8_t)(0x80u | ((cp >> 6) & 0x3Fu)); *d++ = (uint8_t)(0x80u | (cp & 0x3Fu)); }
    else { *d++ = (uint8_t)(0xF0u | (cp >> 18)); *d++ = (uint8_t)(0x80u | ((cp >> 12) & 0x3Fu)); *d++ = (uint8_t)(0x80u | ((cp >> 6) & 0x3Fu)); *d++ = (uint8_t)(0x80u | (cp & 0x3Fu)); }
    return d;
}

// U+03A3 at byte q lower-cases to U+03C2 iff  \p{Cased}\p{Case_Ignorable}* precedes it and no
// \p{Case_Ignorable}*\p{Cased} follows (CPython handle_capital_sigma).  Rare: walks the text in global memory.
static __device__ __noinline__ bool final_sigma(const PretokDev &t, const uint8_t *text, uint64_t n, uint64_t q) {
    uint32_t c = 0, adv; bool found = false;
    uint64_t j = q;
    while (j > 0) {
        uint64_t k = j - 1;
        while (k > 0 && (text[k] & 0xC0u) == 0x80u) --k;
        c = utf8_decode(text + k, n - k < 4 ? (uint32_t)(n - k) : 4u, adv);
        j = k;
        if (!bitmap_bit(t.ignorable, c)) { found = true; break; }
    }
    if (!found || !bitmap_bit(t.cased, c)) return false;
    for (uint64_t k = q + 2; k < n; k += adv) {
        c = utf8_decode(text + k, n - k < 4 ? (uint32_t)(n - k) : 4u, adv);
        if (!bitmap_bit(t.ignorable, c)) return !bitmap_bit(t.cased, c);
    }
    return true;
}

// lower-case mapping of a non-ASCII code point: returns the number of output code points (1 or more) in out[]
__device__ __forceinline__ uint32_t lower_cp(const PretokDev &t, uint32_t cp, const uint8_t *text, uint64_t n, uint64_t q,
                                             uint32_t out[3], uint32_t *status) {
    uint32_t lw = cp < t.n_lower ? __ldg(t.lower + cp) : cp;
    if ((lw & (kLowerMulti | kLowerSigma)) == 0) { out[0] = lw & kLowerCpMask; return 1; }
    if (lw & kLowerSigma) {
        if (!t.cased) { atomicExch(&status[kPtCode], (uint32_t)SWT_ERR_ARG); out[0] = 0x3C3u; return 1; }
        out[0] = final_sigma(t, text, n, q) ? 0x3C2u : 0x3C3u;
        return 1;
    }
    const uint32_t idx = lw & 0xFFFFFu, cnt = min(__ldg(t.multi + idx), 3u);
    for (uint32_t k = 0; k < cnt; ++k) out[k] = __ldg(t.multi + idx + 1 + k);
    return cnt;
}

// ---- analysis of one 32-bit word of text, text[i .. i+4) ------------------------------------------------------------
// All per-byte decisions are taken with SWAR arithmetic on the lane's 32-bit word (flags live in bit 7 of each byte);
// only characters outside ASCII (at most two can start in four bytes) go through a short loop, and the multi-byte
// whitespace characters through a rare exact path that is entered only when one of their lead bytes is in sight.
constexpr uint32_t kHi = 0x80808080u;
__device__ __forceinline__ uint32_t swar_ge(uint32_t x7, uint32_t a) { return (x7 + (0x80u - a) * 0x01010101u) & kHi; }   // per byte: x7 >= a

struct LaneStep {
    uint32_t ns;        // flags: a non-whitespace character starts at this byte
    uint32_t wstart;    // flags: ... and it starts a word
    uint32_t lowered;   // the lane's bytes with ASCII upper case folded
    uint32_t lw[2];     // lower-case table entries of the (up to two) non-ASCII characters, in order
    uint32_t mine;      // (words << 16) | output bytes
};

__device__ __forceinline__ uint32_t decode_at(uint32_t c4) {          // c4: the character's bytes, first byte in bits 0-7
    const uint32_t b0 = c4 & 0xFFu, b1 = (c4 >> 8) & 0x3Fu, b2 = (c4 >> 16) & 0x3Fu, b3 = (c4 >> 24) & 0x3Fu;
    if (b0 < 0xE0u) return ((b0 & 0x1Fu) << 6) | b1;
    if (b0 < 0xF0u) return ((b0 & 0x0Fu) << 12) | (b1 << 6) | b2;
    return ((b0 & 0x07u) << 18) | (b1 << 12) | (b2 << 6) | b3;
}

// rare exact path: multi-byte whitespace in the window.  msp: flags of own bytes that START a multi-byte whitespace
// character; psp: flags of own bytes whose PREVIOUS character is a multi-byte whitespace character.
static __device__ __noinline__ void multibyte_spaces(uint32_t w_prev, uint32_t w_cur, uint32_t w_next, uint32_t &msp, uint32_t &psp) {
    const uint32_t w[3] = {w_prev, w_cur, w_next};
#define B(k) ((w[((k) + 4) >> 2] >> (8 * (((k) + 4) & 3))) & 0xFFu)
    msp = psp = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const uint32_t b0 = B(p), b1 = B(p + 1), b2 = B(p + 2), x = B(p - 1), y = B(p - 2), z = B(p - 3);
        if (is_space_2(b0, b1) || (b0 >= 0xE1u && b0 <= 0xE3u && is_space_3(b0, b1, b2))) msp |= 0x80u << (8 * p);
        if (x >= 0x80u && (is_space_2(y, x) || is_space_3(z, y, x))) psp |= 0x80u << (8 * p);
    }
#undef B
}

// BERT mode: is the character that ENDS just before byte p of the window (a non-ASCII one) punctuation?  window as in B(k)
static __device__ __noinline__ bool prev_char_is_punct(const PretokDev &t, uint32_t w_prev, uint32_t w_cur, int p) {
    const unsigned long long v = ((unsigned long long)w_cur << 32) | w_prev;            // byte k of the window at bits 8 (k + 4)
    int k = p - 1;                                                                       // last byte of the previous character
    while (k > -4 && (((uint32_t)(v >> (8 * (k + 4))) & 0xC0u) == 0x80u)) --k;
    const uint32_t c4 = (uint32_t)(v >> (8 * (k + 4)));
    if ((c4 & 0xFFu) < 0x80u || (c4 & 0xC0u) == 0x80u) return false;
    const uint32_t cp = decode_at(c4);
    return cp < t.n_lower && (__ldg(t.lower + cp) & kLowerPunct) != 0;
}

template <bool kBert>
__device__ __forceinline__ LaneStep analyze(const PretokDev &t, const uint8_t *text, uint64_t n, uint64_t i, uint32_t w_prev,
                                            uint32_t w_cur, uint32_t w_next, uint32_t *status) {
    LaneStep r;
    const uint32_t n_valid = i >= n ? 0u : (n - i >= 4 ? 4u : (uint32_t)(n - i));
    const uint32_t vm = n_valid >= 4 ? kHi : (((1u << (8 * n_valid)) - 1u) & kHi);
    const uint32_t hi = w_cur & kHi, x7 = w_cur & 0x7F7F7F7Fu;
    const uint32_t cont = hi & ~((w_cur << 1) & kHi);                                    // 10xxxxxx
    const uint32_t start = ~cont & vm;
    // ASCII whitespace: str.isspace() = 09-0D, 1C-20; Rust char::is_whitespace (BERT pre-tokenizer) = 09-0D, 20
    uint32_t space = ((swar_ge(x7, 0x09) & ~swar_ge(x7, 0x0E)) | (swar_ge(x7, kBert ? 0x20 : 0x1C) & ~swar_ge(x7, 0x21))) & ~hi;
    const uint32_t upper = swar_ge(x7, 0x41) & ~swar_ge(x7, 0x5B) & ~hi;
    r.lowered = w_cur | (upper >> 2);
    const uint32_t xb = w_prev >> 24;
    const bool xb_space = kBert ? (xb == 0x20u || (xb - 0x09u) < 5u) : is_space_ascii(xb);
    uint32_t prev_space = (space << 8) | ((i == 0 || xb_space) ? 0x80u : 0u);
    // BERT mode: punctuation characters are words of their own (is_bert_punc: ASCII punctuation, or the table bit)
    uint32_t punct = 0, prev_punct = 0;
    if (kBert) {
        punct = ((swar_ge(x7, 0x21) & ~swar_ge(x7, 0x30)) | (swar_ge(x7, 0x3A) & ~swar_ge(x7, 0x41)) | (swar_ge(x7, 0x5B) & ~swar_ge(x7, 0x61)) |
                 (swar_ge(x7, 0x7B) & ~swar_ge(x7, 0x7F))) & ~hi;
        const bool xb_punct = (xb - 0x21u) < 15u || (xb - 0x3Au) < 7u || (xb - 0x5Bu) < 6u || (xb - 0x7Bu) < 4u;
        prev_punct = (punct << 8) | (xb_punct ? 0x80u : 0u);
    }
    // lead bytes of the multi-byte whitespace characters: C2, E1, E2, E3 (within three bytes before, or in, the lane's word)
    const uint32_t p7 = w_prev & 0x7F7F7F7Fu;
    const uint32_t lead_cur = hi & ((swar_ge(x7, 0x42) & ~swar_ge(x7, 0x43)) | (swar_ge(x7, 0x61) & ~swar_ge(x7, 0x64)));
    const uint32_t lead_prev = (w_prev & 0x80808000u) & ((swar_ge(p7, 0x42) & ~swar_ge(p7, 0x43)) | (swar_ge(p7, 0x61) & ~swar_ge(p7, 0x64)));
    if (lead_cur | lead_prev) {
        uint32_t msp, psp;
        multibyte_spaces(w_prev, w_cur, w_next, msp, psp);
        space |= msp; prev_space |= psp;
    }
    r.ns = start & ~space;
    if (kBert) {
        // character starts that follow a non-ASCII character: that character may be punctuation
        uint32_t after_hi = r.ns & ((hi << 8) | (xb >= 0x80u ? 0x80u : 0u));
        while (after_hi) {
            const uint32_t bit = __ffs(after_hi) - 1; after_hi &= after_hi - 1;
            if (prev_char_is_punct(t, w_prev, w_cur, (int)((bit - 7) >> 3))) prev_punct |= 1u << bit;
        }
    }
    uint32_t bytes = __popc(r.ns & ~hi);
    r.lw[0] = r.lw[1] = 0;
    uint32_t nas = r.ns & hi;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (nas) {
            const uint32_t bit = __ffs(nas) - 1; nas &= nas - 1;                          // bit = 8 p + 7
            const uint32_t cp = decode_at(__funnelshift_r(w_cur, w_next, bit - 7));
            const uint32_t lw = cp < t.n_lower ? __ldg(t.lower + cp) : cp;
            r.lw[k] = lw;
            if (kBert && (lw & kLowerPunct)) punct |= 1u << bit;
            if ((lw & (kLowerMulti | kLowerSigma)) == 0) bytes += utf8_len_of(lw & kLowerCpMask);
            else {
                uint32_t out[3];
                const uint32_t cnt = lower_cp(t, cp, text, n, i + ((bit - 7) >> 3), out, status);
                for (uint32_t c = 0; c < cnt; ++c) bytes += utf8_len_of(out[c]);
            }
        }
    }
    r.wstart = r.ns & (prev_space | prev_punct | punct);
    r.mine = ((uint32_t)__popc(r.wstart) << 16) | bytes;
    return r;
}

// writes the lane's lowered bytes at arena + byte_pos and its word offsets at word_off + word_pos
__device__ __forceinline__ void emit(const PretokDev &t, const uint8_t *text, uint64_t n, uint64_t i, uint32_t w_cur, uint32_t w_next,
                                     const LaneStep &r, uint8_t *arena, uint64_t byte_pos, uint32_t *word_off, uint32_t *word_src, uint32_t word_pos,
                                     uint32_t *status) {
    uint8_t *dst = arena + byte_pos;
    uint32_t used = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const uint32_t flag = 0x80u << (8 * p);
        if (!(r.ns & flag)) continue;
        if (r.wstart & flag) {
            if (word_src) word_src[word_pos] = (uint32_t)(i + p);             // where the word starts in the text
            word_off[word_pos++] = (uint32_t)(dst - arena);
        }
        const uint32_t b = (r.lowered >> (8 * p)) & 0xFFu;
        if (b < 0x80u) { *dst++ = (uint8_t)b; continue; }
        if (used == 2) continue;                            // only reachable with malformed UTF-8: never write more than was counted
        const uint32_t lw = used ? r.lw[1] : r.lw[0];
        ++used;
        if ((lw & (kLowerMulti | kLowerSigma)) == 0) dst = put_utf8(dst, lw & kLowerCpMask);
        else {
            uint32_t out[3];
            const uint32_t cnt = lower_cp(t, decode_at(__funnelshift_r(w_cur, w_next, 8 * p)), text, n, i + p, out, status);
            for (uint32_t c = 0; c < cnt; ++c) dst = put_utf8(dst, out[c]);
        }
    }
}

// text loads: read-only path for text in global memory (kLdg), plain loads when the tile functions run on a copy in shared memory
template <bool kLdg> __device__ __forceinline__ uint32_t ld_text32(const uint32_t *p) { if constexpr (kLdg) return __ldg(p); else return *p; }
template <bool kLdg> __device__ __forceinline__ uint4 ld_text128(const uint4 *p) { if constexpr (kLdg) return __ldg(p); else return *p; }

// count pass over one 4 KiB tile by one warp: a lane owns 16 consecutive bytes per step (one 128-bit load, 512 bytes per warp step);
// the words before and after them come from the neighbour lanes.  Returns (words << 32) | bytes of the tile on every lane.
template <bool kBert, bool kLdg = true>
__device__ __forceinline__ unsigned long long count_tile(const PretokDev &t, const uint8_t *__restrict__ text, uint64_t n, uint32_t tile, uint32_t *status) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t *t32 = reinterpret_cast<const uint32_t *>(text);
    const uint64_t n_words32 = (n + 3) >> 2, t0 = (uint64_t)tile * kTileBytes;
    uint32_t tile_words = 0, tile_bytes = 0;
    for (uint32_t s = 0; s < kTileBytes / 512; ++s) {
        const uint64_t step0 = t0 + (uint64_t)s * 512;
        if (step0 >= n) break;                                              // warp-uniform
        const uint64_t i = step0 + lane * 16, wi = i >> 2;
        uint4 q;
        if (wi + 4 <= n_words32) q = ld_text128<kLdg>(reinterpret_cast<const uint4 *>(t32 + wi));
        else {
            q.x = wi < n_words32 ? ld_text32<kLdg>(t32 + wi) : 0u; q.y = wi + 1 < n_words32 ? ld_text32<kLdg>(t32 + wi + 1) : 0u;
            q.z = wi + 2 < n_words32 ? ld_text32<kLdg>(t32 + wi + 2) : 0u; q.w = 0u;
        }
        uint32_t w_prev = __shfl_up_sync(0xffffffffu, q.w, 1), w_next = __shfl_down_sync(0xffffffffu, q.x, 1);
        if (lane == 0) w_prev = wi >= 1 ? ld_text32<kLdg>(t32 + wi - 1) : 0u;
        if (lane == 31) w_next = wi + 4 < n_words32 ? ld_text32<kLdg>(t32 + wi + 4) : 0u;
        const uint32_t mine = analyze<kBert>(t, text, n, i, w_prev, q.x, q.y, status).mine + analyze<kBert>(t, text, n, i + 4, q.x, q.y, q.z, status).mine +
                              analyze<kBert>(t, text, n, i + 8, q.y, q.z, q.w, status).mine + analyze<kBert>(t, text, n, i + 12, q.z, q.w, w_next, status).mine;
        tile_words += mine >> 16; tile_bytes += mine & 0xFFFFu;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        tile_words += __shfl_xor_sync(0xffffffffu, tile_words, d);
        tile_bytes += __shfl_xor_sync(0xffffffffu, tile_bytes, d);
    }
    return ((unsigned long long)tile_words << 32) | tile_bytes;
}

// write pass over one tile by one warp, from the tile's output position `base` = (first word << 32) | first byte: a lane owns 4
// consecutive bytes per step (128 bytes per warp step), so that the byte stores of a warp land in one contiguous run; the loads of
// four steps are issued together
template <bool kBert, bool kLdg = true>
__device__ __forceinline__ void write_tile(const PretokDev &t, const uint8_t *__restrict__ text, uint64_t n, uint32_t tile, unsigned long long base,
                                           uint8_t *__restrict__ arena, uint32_t *__restrict__ word_off, uint32_t *__restrict__ word_src, uint32_t *status) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t *t32 = reinterpret_cast<const uint32_t *>(text);
    const uint64_t n_words32 = (n + 3) >> 2, t0 = (uint64_t)tile * kTileBytes;
    uint64_t byte_pos = base & 0xFFFFFFFFull; uint32_t word_pos = (uint32_t)(base >> 32);
    for (uint32_t s4 = 0; s4 < kTileBytes / 512; ++s4) {
        const uint64_t blk0 = t0 + (uint64_t)s4 * 512;
        if (blk0 >= n) break;                                               // warp-uniform
        uint32_t wq[6];                                                     // words lane-1 .. of the four steps: wq[k+1] = step k
        const uint64_t wi0 = (blk0 >> 2) + lane;
#pragma unroll
        for (int k = 0; k < 4; ++k) wq[k + 1] = wi0 + 32 * k < n_words32 ? ld_text32<kLdg>(t32 + wi0 + 32 * k) : 0u;
        wq[0] = (lane == 0 && wi0 >= 1) ? ld_text32<kLdg>(t32 + wi0 - 1) : 0u;
        wq[5] = (lane == 31 && wi0 + 97 < n_words32) ? ld_text32<kLdg>(t32 + wi0 + 97) : 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint64_t step0 = blk0 + 128 * k;
            if (step0 >= n) break;                                          // warp-uniform
            const uint64_t i = step0 + lane * 4;
            const uint32_t w_cur = wq[k + 1];
            uint32_t w_prev = __shfl_up_sync(0xffffffffu, w_cur, 1), w_next = __shfl_down_sync(0xffffffffu, w_cur, 1);
            // lane 0's previous word is lane 31's word of the previous step; lane 31's next word is lane 0's of the next
            const uint32_t from31 = __shfl_sync(0xffffffffu, wq[k], 31), from0 = __shfl_sync(0xffffffffu, wq[k + 2 > 5 ? 5 : k + 2], 0);
            if (lane == 0) w_prev = k == 0 ? wq[0] : from31;
            if (lane == 31) w_next = k == 3 ? wq[5] : from0;
            const LaneStep r = analyze<kBert>(t, text, n, i, w_prev, w_cur, w_next, status);
            uint32_t incl = r.mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += v; }
            const uint32_t excl = incl - r.mine, total = __shfl_sync(0xffffffffu, incl, 31);
            if (r.ns) emit(t, text, n, i, w_cur, w_next, r, arena, byte_pos + (excl & 0xFFFFu), word_off, word_src, word_pos + (excl >> 16), status);
            byte_pos += total & 0xFFFFu; word_pos += total >> 16;
        }
    }
}

inline PretokDev dev_view(const swt_pretok *p) { return PretokDev{p->d_lower, p->n_lower, p->d_multi, p->d_cased, p->d_ignorable}; }

}  // namespace pt
}  // namespace swt
