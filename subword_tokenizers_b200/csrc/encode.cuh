// encode.cuh -- the tile kernels shared by all encoders (HP-1 FastBPE, HP-2 FastWP, and the NaiveBPE / NaiveWP encoders).
//
// Data flow of one encode call (north-star subsystem 1: packed word-offset/byte arena).  Tile = 64 consecutive words per
// WARP, two rows of 32 (word w_tile + 32 j + lane on lane `lane`); tiles assigned round-robin to a persistent grid; warps never
// synchronise with each other.
//
//   pass 1  encode_count_kernel   per word: look the word up in the word-type memo (one scattered 16-byte load); rare cases -- first
//                                 occurrences, words of 16..32 bytes, 14+ tokens -- go to the warp's pending queue and are resolved
//                                 32 at a time (probe at L2, encode through the rank table / trie, publish).  Writes one packed
//                                 u32 record per word (kind, token count, memo slot) and the tile's token total.
//           split form            (Enc::kSplitCount, FastBPE): warm-up by the kernel above, then encode_count_leaf_kernel (probe
//                                 only, no encoder / calls / stack) + encode_resolve_kernel per chunk of the stream
//           encode_long_count     words longer than 32 bytes, one warp per flagged tile
//   scan    two tiny kernels      exclusive scan of the tile totals (in-group prefixes + group bases)
//   pass 2  encode_emit_kernel    per word: one scattered load of ids16[slot], ids expanded into the warp's shared-memory buffer at
//                                 their tile-local position; the tile's ids leave shared memory as ONE bulk copy (cp.async.bulk
//                                 shared::cta -> global, issued by one lane: no LSU wavefronts) at their final position, u32 token
//                                 offsets as coalesced stores.  encode_long_emit: the ids of long words.
//   small   tokenize_small_kernel one CTA pre-tokenizes and encodes one short text (the per-line tokenize() call)
//
// Why two passes: a single pass needs every tile's global token base while the tile is in flight.  With warp-sized tiles a decoupled
// look-back walked hundreds of in-flight predecessors; with CTA-sized tiles every tile waits for ALL earlier tiles, and a tile that
// holds a first occurrence (a 10-25 us trie walk) stalls every successor -- head-of-line blocking.  The two passes cost 8 extra bytes
// of HBM traffic per word and a second scattered load per word, and have no inter-tile dependency at all.
//
// Word-type memo: encode_word is a pure function of the word and word streams are
// Zipf-distributed, so every launch keeps a hash table  word bytes -> token ids  in its workspace.  The first
// thread that meets a word type claims a slot with ONE 128-bit CAS (key = first 15 bytes + length; the
// remaining bytes of longer words are stored beside the entry and compared, so a hit is always exact), encodes
// the word and publishes the ids with a release; later occurrences copy the ids instead of re-walking
// the rank table / trie.  The memo is rebuilt from empty by every launch (nothing is carried over between
// calls), readers never wait (a slot that is claimed but not yet published is simply recomputed), and words
// that do not fit (longer than 32 bytes, table full) take the direct path.
// Every input byte is still read and every output id still written by every launch.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "pretok.cuh"

namespace swt {

constexpr int kWarps = 8;              // warps per CTA; every warp owns its own tile
constexpr int kThreads = kWarps * 32;
#ifndef SWT_COUNT_CTAS
#define SWT_COUNT_CTAS 4
#endif
#ifndef SWT_EMIT_CTAS
#define SWT_EMIT_CTAS 4
#endif
constexpr int kCtasPerSm = SWT_COUNT_CTAS;        // resident CTAs per SM the count / emit kernels are compiled for (register budget)
constexpr int kEmitCtasPerSm = SWT_EMIT_CTAS;
constexpr int kWordsPerThread = 2;
constexpr int kTileWords = 32 * kWordsPerThread;         // 64 words per (warp) tile
constexpr int kShortBytes = 32;        // words up to this many bytes are encoded by one thread (and memoised)
constexpr int kCompactTokens = 640;    // tile token totals up to this are assembled in smem before the store
constexpr int kMemoTokens = 44;        // >= kShortBytes: every word of up to 32 bytes fits (ids <= bytes)
constexpr int kMemoProbes = 8;

// status words written by the encode kernels
enum { kStatusCode = 0, kStatusTokens = 1, kStatusH6 = 2, kStatusTokensHi = 3, kStatusMemoTypes = 4, kStatusSlowWords = 5,
       kStatusWarmSlow = 6 };   // slow-path words of the last warm-up tiles: gates the split count pass (see launch_encode_tiles)
// split count pass: the stream has MANY word types when more than 1/8 of the words of the last warm-up tiles took the slow path
// (counted over the LAST QUARTER of the warm-up tiles only: the first waves of tiles miss whatever the stream, nothing is published yet)
__device__ __forceinline__ bool many_types_stream(const uint32_t *status, uint32_t warm_tiles) {
    return (uint64_t)status[kStatusWarmSlow] * 8u > (uint64_t)(warm_tiles / 4u) * 64u;
}

// Word-type memo, structure of arrays (one slot index addresses all three):
//   keys[slot]   16 B  the CAS key: word bytes 0..7 | bytes 8..14, bits 56-59 = length (0 for words > 15 bytes), bits 60-63 = "pub":
//                      0 = claimed, not published; 1..13 = n_tokens and ids16[slot] is valid; 14 = ids16[slot] valid with 14..16
//                      tokens (count in ext[slot].meta); 15 = published in ext[slot] (ids that do not fit 16 bits, more than 16
//                      tokens, H6 events, or "not cacheable").  The ONLY array a call clears.
//   ids16[slot]  32 B  up to 16 ids as 16-bit values (Enc::narrow16 / Enc::expand16), read by the emit pass
//   ext[slot]    32 B  tail of words of 16..32 bytes (compared on every hit, so a hit is exact), meta and the position of the 32-bit
//                      id list in the tok32 arena (bump allocated)
// Round 1 kept one 256-byte record per word type: a memo for millions of types then spans gigabytes and the hot entries of a
// Zipf stream each sit in their own 2 MB page (TLB reach: 2^22 entries cost the bench stream 30 %).  With 16-byte keys the
// count pass probes a table of 64 MB at 2^22 slots, two keys per sector, and nothing else has to be cleared.
struct alignas(32) MemoExt {
    uint32_t meta;                      // (n_tokens + 1) | (h6 << 8); 0xFFFFFFFF = not cacheable
    uint32_t tail_last;                 // words > 15 bytes: byte 31 | length << 8
    unsigned long long tail_a, tail_b;  // words > 15 bytes: bytes 15..22 | 23..30
    uint32_t tok32_off;                 // first id of the 32-bit list in EncodeWorkspace::tok32
    uint32_t pad;
};
constexpr uint32_t kMemoSlotBits = 23, kMemoSlotMask = (1u << kMemoSlotBits) - 1u;     // slot field of the per-word record
static_assert(sizeof(MemoExt) == 32, "MemoExt must be 32 bytes");
constexpr unsigned long long kPubMask = 0xFull << 60;
constexpr uint32_t kPubExt = 15;          // pub nibble: look in ext[slot]
constexpr uint32_t kPubWide = 14;         // pub nibble: ids16[slot] valid, 14..16 tokens (the count is in ext[slot].meta unless the length pins it)
constexpr uint32_t kNarrowMaxTokens = 16; // pub nibble 1..13 = n_tokens (a memoised word has at least one token), 14 = kPubWide

struct MemoKey { unsigned long long lo, hi, tail_a, tail_b; uint32_t tail_last; uint32_t nbytes; };

struct EncodeWorkspace {
    unsigned long long *long_cursor; // [0]: allocation cursor into long_scratch (u32 units); [1]: bump cursor into tok32
    uint4 *keys; uint32_t memo_mask; // memo_mask == 0: memo disabled
    uint4 *ids16;                    // 2 x uint4 per slot
    MemoExt *ext;
    uint32_t *tok32; uint32_t tok32_cap;
    uint32_t *packed;                // n_words: per-word record handed from the count pass to the emit pass
    uint32_t *tile_total;            // n_tiles: tokens per tile, then (after the scan) the in-group exclusive prefix
    uint2 *tile_pend;                // n_tiles: per row, the words the leaf count kernel left to the resolve kernel (bit = lane)
    unsigned long long *group_base;  // n_groups: tokens per group of 1024 tiles, then the group's global token offset
    uint32_t *long_tiles;            // bitmap over the tiles: the tile holds a word longer than kShortBytes (rare; cleared per call)
    uint32_t *long_scratch;          // BPE: symbol ping-pong buffers of long words; WP: segment records of long chunks
    uint64_t long_scratch_elems;
    uint32_t n_tiles;
    uint32_t flags;                  // (unused)
};

struct Tuning {                     // process-wide knobs for experiments (swt_tune); defaults are the measured best
    int memo_max_log2 = 22;         // cap of the memo size (10..23)
    int memo_off = 0;               // 1: no memo -- every word takes the direct path (reported as *_direct rates)
    int bulk_store = 1;             // (kept for old experiment scripts: the copy-out is always cp.async.bulk now)
    int timing = 0;                 // per-kernel CUDA-event times of every encode call on stderr (synchronises)
    int warp_words = 3;             // FastBPE: up to this many missed words are encoded by the whole warp, one after the other
    int bpe_queue = 1;              // FastBPE: memo misses go through the warp's pending queue (0: resolved inside their tile)
    int split_count = 1;            // count pass = warm-up + chunks of (leaf count kernel, resolve kernel); 0: one kernel with the slow path inside
    int split_chunks = 2;           // ... number of chunks
    int split_warm_tiles = 16384;   // ... tiles of the warm-up (1 M words); the split form is used for calls of more than 4x as many tiles
};
extern Tuning g_tune;

size_t encode_workspace_layout(uint32_t n_words, uint64_t long_bytes, void *base, EncodeWorkspace *ws);
int encode_grid(const void *kernel, int block, size_t dyn_smem);

#ifdef __CUDACC__
enum { kMemoHit = 0, kMemoClaimed = 1, kMemoMiss = 2, kMemoPending = 3 };

__device__ __forceinline__ void cas128(uint4 *e, unsigned long long lo, unsigned long long hi,
                                       unsigned long long &old_lo, unsigned long long &old_hi) {
    asm volatile("{\n\t.reg .b128 c, v, o;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 v, {%4, %5};\n\t"
                 "atom.global.cas.b128 o, [%6], c, v;\n\tmov.b128 {%0, %1}, o;\n\t}"
                 : "=l"(old_lo), "=l"(old_hi) : "l"(0ull), "l"(0ull), "l"(lo), "l"(hi), "l"(e) : "memory");
}
// the memo is written during the launch: the slow path reads it at L2 (never through the non-coherent L1)
__device__ __forceinline__ void ld_cg_u64x2(const void *p, unsigned long long &a, unsigned long long &b) {
    asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ uint4 ld_cg_u32x4(const void *p) {
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// L1-cached variants for the FAST path.  Hot word types (Zipf) then hit the 100+ KB L1 instead of going to L2.  This is
// safe although the memo is written during the launch: a key never changes once set and its pub nibble goes 0 -> final exactly
// once, after the ids / ext record were written and fenced (release); L1 fills are whole 32-byte sectors.  A stale L1 sector can
// therefore only show "empty" or "not published yet", which sends the word to the slow path, and the slow path re-probes
// at L2 (ld.cg).  ids16[] / ext[] of a slot are only ever read after a non-zero pub nibble was seen, i.e. after publication;
// nobody reads them before (so no stale copy of them can sit in an L1), and the emit pass is a later launch.
__device__ __forceinline__ void ld_ca_u64x2(const void *p, unsigned long long &a, unsigned long long &b) {
    asm volatile("ld.global.ca.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ uint4 ld_ca_u32x4(const void *p) {
    uint4 v;
    asm volatile("ld.global.ca.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// (no acquire loads anywhere: ld.acquire.gpu compiles to LDG + CCTL.IVALL, a full L1 invalidate per probe.  Ordering comes from the
// writer's fence -- ids are performed at L2 before the pub nibble -- plus the reader's control dependency.)

// predicated (branch-free) read-only load: 0 when the predicate is false
__device__ __forceinline__ uint32_t ldg_u32_if(const uint32_t *p, bool pred) {
    uint32_t v = 0;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p ld.global.nc.u32 %0, [%1];\n\t}" : "+r"(v) : "l"(p), "r"((uint32_t)pred));
    return v;
}
// mask of the lowest min(max(n, 0), 4) bytes of a 32-bit word
__device__ __forceinline__ uint32_t low_bytes_mask(int n) { return __funnelshift_lc(0xFFFFFFFFu, 0u, (uint32_t)max(n, 0) * 8u); }

// Key of a word of 1..32 bytes: the first 15 bytes + the length form the 128-bit CAS key, bytes 15..31 the tail.
// Reads aligned 8-byte words when that stays inside the arena.
__device__ __forceinline__ void memo_key(const uint8_t *arena, uint32_t b0, uint32_t nbytes, uint32_t arena_end, MemoKey &k) {
    const uint8_t *p = arena + b0;
    unsigned long long w0 = 0, w1 = 0, w2 = 0, w3 = 0;
    if ((uint64_t)b0 + 40 <= arena_end) {
        const uintptr_t a = (uintptr_t)p;
        const unsigned long long *q = (const unsigned long long *)(a & ~(uintptr_t)7);
        const uint32_t sh = (uint32_t)(a & 7) * 8;
        const unsigned long long r0 = __ldg(q), r1 = __ldg(q + 1);
        const unsigned long long r2 = nbytes + (sh >> 3) > 16 ? __ldg(q + 2) : 0ull;
        const unsigned long long r3 = nbytes + (sh >> 3) > 24 ? __ldg(q + 3) : 0ull;
        const unsigned long long r4 = nbytes + (sh >> 3) > 32 ? __ldg(q + 4) : 0ull;
        if (sh) {
            w0 = (r0 >> sh) | (r1 << (64 - sh)); w1 = (r1 >> sh) | (r2 << (64 - sh));
            w2 = (r2 >> sh) | (r3 << (64 - sh)); w3 = (r3 >> sh) | (r4 << (64 - sh));
        } else { w0 = r0; w1 = r1; w2 = r2; w3 = r3; }
    } else {
        for (uint32_t i = 0; i < nbytes; ++i) {
            const unsigned long long c = (unsigned long long)p[i] << (8 * (i & 7));
            if (i < 8) w0 |= c; else if (i < 16) w1 |= c; else if (i < 24) w2 |= c; else w3 |= c;
        }
    }
    // zero the bytes at and beyond nbytes
    if (nbytes < 8) { w0 &= (1ull << (8 * nbytes)) - 1; w1 = w2 = w3 = 0; }
    else if (nbytes < 16) { w1 = nbytes == 8 ? 0ull : w1 & ((1ull << (8 * (nbytes - 8))) - 1); w2 = w3 = 0; }
    else if (nbytes < 24) { w2 = nbytes == 16 ? 0ull : w2 & ((1ull << (8 * (nbytes - 16))) - 1); w3 = 0; }
    else if (nbytes < 32) { w3 = nbytes == 24 ? 0ull : w3 & ((1ull << (8 * (nbytes - 24))) - 1); }
    k.lo = w0;
    k.hi = (w1 & ((1ull << 56) - 1)) | ((unsigned long long)(nbytes <= 15 ? nbytes : 0u) << 56);
    k.tail_a = (w1 >> 56) | (w2 << 8);
    k.tail_b = (w2 >> 56) | (w3 << 8);
    k.tail_last = (uint32_t)(w3 >> 56) | (nbytes << 8);
    k.nbytes = nbytes;
}

// slot hash of a key: a multilinear form of the four key words (four IMADs) and one avalanche round -- half the
// instructions of the murmur-style mixer of round 1, on the path every word takes
__device__ __forceinline__ uint32_t memo_hash4(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
    uint32_t h = w0 * 0x9E3779B1u + w1 * 0x85EBCA77u + w2 * 0xC2B2AE3Du + w3 * 0x27D4EB2Fu + 0x165667B1u;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 13;
    return h;
}
__device__ __forceinline__ uint32_t memo_hash(unsigned long long lo, unsigned long long hi) {
    return memo_hash4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32) & 0x0FFFFFFFu);
}

// Probes the memo at L2 (slow path).  kMemoHit: slot/meta describe a published entry for exactly this word (`narrow`: its ids
// live in ids16[] and meta is the pub nibble, else meta is MemoExt::meta and the ids are in the tok32 arena).
// kMemoClaimed: this thread now owns `slot` and must call memo_publish after encoding.  kMemoPending: another thread owns the
// slot of exactly this word (words of up to 15 bytes: the key is the whole word) and will have published it by the time the
// emit pass runs -- encode for the count, but let the emit pass read the ids from `slot`.  kMemoMiss: encode directly, publish nothing.
static __device__ __noinline__ int memo_probe(const EncodeWorkspace &ws, const MemoKey &key, uint32_t &slot, uint32_t &ntok_out, uint32_t &h6_out, bool &narrow) {
    uint32_t h = memo_hash(key.lo, key.hi) & ws.memo_mask;
    for (int probe = 0; probe < kMemoProbes; ++probe, h = (h + 1) & ws.memo_mask) {
        uint4 *e = ws.keys + h;
        unsigned long long klo, khi;
        ld_cg_u64x2(e, klo, khi);
        if (klo == 0 && khi == 0) {
            cas128(e, key.lo, key.hi, klo, khi);
            if (klo == 0 && khi == 0) { slot = h; return kMemoClaimed; }
            // lost the race: (klo, khi) is what the winner stored (its pub nibble may already be set)
        }
        if (klo != key.lo || (khi & ~kPubMask) != key.hi) continue;
        const uint32_t pub = (uint32_t)(khi >> 60);
        if (pub == 0) {                                                   // claimed by another thread, not published yet
            if (key.nbytes <= 15) { slot = h; return kMemoPending; }
            return kMemoMiss;                                             // long word: its tail is not comparable yet
        }
        uint4 x = make_uint4(0, 0, 0, 0);
        if (key.nbytes > 15 || pub >= kPubWide) x = ld_cg_u32x4(&ws.ext[h]);   // meta, tail_last, tail_a
        if (key.nbytes > 15) {                                            // same 15-byte prefix: check the rest and the length
            const unsigned long long ta = (unsigned long long)x.z | ((unsigned long long)x.w << 32);
            unsigned long long tb, dummy;
            ld_cg_u64x2(&ws.ext[h].tail_b, tb, dummy);
            if (ta != key.tail_a || tb != key.tail_b || x.y != key.tail_last) continue;
        }
        narrow = pub != kPubExt;
        if (!narrow && x.x == 0xFFFFFFFFu) return kMemoMiss;              // not cacheable
        slot = h;
        ntok_out = pub < kPubWide ? pub : (x.x & 0xFFu) - 1u;
        h6_out = narrow ? 0u : x.x >> 8;
        return kMemoHit;
    }
    return kMemoMiss;
}

// per-word record between the two passes, packed into 32 bits:
//   [31:29] kind; Hit (32-bit ids in tok32) / Hit16 (ids16[] valid): [28:23] n_tokens, [22:0] memo slot; Recompute: [28:23] n_tokens;
//   WP long, walked by its lane: [28:0] n_tokens; long word with scratch (BPE; WP segment records): [28:0] scratch granule (16 u32),
//   header word 0 of the granule holds n_tokens
// (numbered so that the emit pass tests the common case with one compare: records >= kWordHit << 29 are the unusual ones)
enum : uint32_t { kWordNone = 0u, kWordHit16 = 1u, kWordHit = 2u, kWordRecompute = 3u, kWordLong = 5u, kWordLongB = 6u, kWordLongSeg = 7u };
constexpr uint32_t kLongHeader = 16;      // u32 words reserved in front of the scratch of a long word
constexpr uint32_t kGroupTiles = 1024;    // tiles per scan group

template <class Enc>
__device__ __forceinline__ bool ids_are_narrow(const uint32_t *buf, uint32_t ntok, uint32_t h6) {
    bool nar = ntok >= 1 && ntok <= kNarrowMaxTokens && h6 == 0;
    uint32_t dummy;
    for (uint32_t k = 0; k < ntok && nar; ++k) nar = Enc::narrow16(buf[k], k, dummy);
    return nar;
}

// Publishes the ids of a claimed slot.  Returns the record kind under which this occurrence is served: kWordHit16 (ids16[]),
// kWordHit (tok32 arena) or kWordRecompute (not cacheable: the arena is full).
template <class Enc>
__device__ __forceinline__ uint32_t memo_publish(const EncodeWorkspace &ws, uint32_t slot, const MemoKey &key, const uint32_t *buf,
                                                 uint32_t ntok, uint32_t h6, uint32_t *status) {
    MemoExt *x = ws.ext + slot;
    if (key.nbytes > 15) { x->tail_a = key.tail_a; x->tail_b = key.tail_b; x->tail_last = key.tail_last; }
    uint32_t pub, kind;
    if (ntok >= 1 && ids_are_narrow<Enc>(buf, ntok, h6)) {
        uint32_t v[16];
#pragma unroll
        for (uint32_t k = 0; k < 16; ++k) { v[k] = 0; if (k < ntok) Enc::narrow16(buf[k], k, v[k]); }
        ws.ids16[2 * (size_t)slot] = make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
        if (ntok > 8) ws.ids16[2 * (size_t)slot + 1] = make_uint4(v[8] | (v[9] << 16), v[10] | (v[11] << 16), v[12] | (v[13] << 16), v[14] | (v[15] << 16));
        if (ntok >= kPubWide) { x->meta = ntok + 1; pub = kPubWide; } else pub = ntok;
        kind = kWordHit16;
    } else {
        const unsigned long long off = ntok <= (uint32_t)kMemoTokens && h6 <= 0xFFFFFFu ? atomicAdd(ws.long_cursor + 1, (unsigned long long)ntok) : ~0ull;
        if (off != ~0ull && off + ntok <= ws.tok32_cap) {
            for (uint32_t k = 0; k < ntok; ++k) ws.tok32[off + k] = buf[k];
            x->tok32_off = (uint32_t)off; x->meta = (ntok + 1) | (h6 << 8); kind = kWordHit;
        } else { x->meta = 0xFFFFFFFFu; kind = kWordRecompute; }
        pub = kPubExt;
    }
    __threadfence();                                                      // release: everything above is performed before the pub nibble
    atomicOr(reinterpret_cast<uint32_t *>(ws.keys + slot) + 3, pub << 28);
    atomicAdd(&status[kStatusMemoTypes], 1u);
    return kind;
}

struct SlowResult { uint32_t kind, ntok, slot, h6; };

// pass 1 slow path (first probe did not hit): full probe at L2, then direct encode + publish.  Out of line so that
// the unrolled fast path stays small.
template <class Enc>
__device__ __noinline__ SlowResult resolve_slow(const Enc &enc, const typename Enc::Stage *sg, const EncodeWorkspace &ws, const uint8_t *arena,
                                                uint32_t b0, uint32_t nbytes, uint32_t arena_end, uint32_t *status) {
    uint32_t buf[kShortBytes];                                   // scratch for one directly encoded word
    SlowResult r; r.kind = kWordRecompute; r.ntok = 0; r.slot = 0; r.h6 = 0;
    int m = kMemoMiss; uint32_t hit_ntok = 0, hit_h6 = 0; MemoKey key; bool narrow = false;
    if (ws.memo_mask && nbytes >= 1) { memo_key(arena, b0, nbytes, arena_end, key); m = memo_probe(ws, key, r.slot, hit_ntok, hit_h6, narrow); }
    if (m == kMemoHit) {
        r.kind = narrow ? kWordHit16 : kWordHit; r.ntok = hit_ntok; r.h6 = hit_h6;
        return r;
    }
    r.ntok = enc.encode_short(sg, arena + b0, nbytes, buf, r.h6);
    if (m == kMemoClaimed) r.kind = memo_publish<Enc>(ws, r.slot, key, buf, r.ntok, r.h6, status);
    // Pending: the owner publishes the same ids in the same form (memo_publish is a function of the ids); an id list that goes to the
    // tok32 arena may turn out not cacheable, which the emit pass sees in ext[slot].meta and answers by re-encoding
    else if (m == kMemoPending) r.kind = ids_are_narrow<Enc>(buf, r.ntok, r.h6) ? kWordHit16 : kWordHit;
    return r;
}
// The same for ONE word by the whole warp (Enc::kWarpShort): lane 0 probes (and claims), all lanes encode (encode_short_warp leaves
// the ids in `ids`, 32 words of shared memory), lane 0 publishes.  The result is returned on every lane.
template <class Enc>
__device__ __noinline__ SlowResult resolve_slow_warp(const Enc &enc, const typename Enc::Stage *sg, const EncodeWorkspace &ws, const uint8_t *arena,
                                                     uint32_t b0, uint32_t nbytes, uint32_t arena_end, uint32_t *status, uint32_t *ids) {
    const uint32_t lane = threadIdx.x & 31;
    int pm = kMemoMiss; uint32_t pslot = 0, pntok = 0, ph6 = 0; bool pnarrow = false; MemoKey key;
    if (lane == 0 && ws.memo_mask && nbytes >= 1) { memo_key(arena, b0, nbytes, arena_end, key); pm = memo_probe(ws, key, pslot, pntok, ph6, pnarrow); }
    pm = __shfl_sync(0xffffffffu, pm, 0);
    SlowResult r; r.kind = kWordRecompute; r.ntok = 0; r.slot = pslot; r.h6 = 0;
    if (pm == kMemoHit) {
        r.kind = pnarrow ? kWordHit16 : kWordHit; r.ntok = pntok;
    } else {
        const uint32_t n = enc.encode_short_warp(sg, arena + b0, nbytes, ids);       // ids[0..n)
        __syncwarp();
        r.ntok = n;
        if (lane == 0) {
            if (pm == kMemoClaimed) r.kind = memo_publish<Enc>(ws, pslot, key, ids, n, 0u, status);
            else if (pm == kMemoPending) r.kind = ids_are_narrow<Enc>(ids, n, 0u) ? kWordHit16 : kWordHit;
        }
        __syncwarp();
    }
    r.kind = __shfl_sync(0xffffffffu, r.kind, 0); r.ntok = __shfl_sync(0xffffffffu, r.ntok, 0); r.slot = __shfl_sync(0xffffffffu, r.slot, 0);
    return r;
}

// The slow words of one tile row (ballot m; this lane's word is b0 / nbytes when its bit is set), resolved now: up to `warp_words` of
// them one after the other by the whole warp (Enc::kWarpShort), more of them one word per lane.  Returns this lane's own result
// (h6: the events this lane counted while resolving, for whichever word).
template <class Enc>
__device__ __noinline__ SlowResult resolve_slow_tile(const Enc &enc, const typename Enc::Stage *sg, const EncodeWorkspace &ws, const uint8_t *arena,
                                                     uint32_t arena_end, uint32_t *status, uint32_t *ids, uint32_t m, uint32_t b0, uint32_t nbytes,
                                                     uint32_t warp_words) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t n_slow = __popc(m);
    SlowResult mine; mine.kind = kWordNone; mine.ntok = 0; mine.slot = 0; mine.h6 = 0;
    if constexpr (Enc::kWarpShort) {
        if (n_slow <= warp_words) {
            for (uint32_t mm = m; mm;) {
                const uint32_t owner = __ffs(mm) - 1; mm &= mm - 1;
                const uint32_t sb0 = __shfl_sync(0xffffffffu, b0, owner), snb = __shfl_sync(0xffffffffu, nbytes, owner);
                const SlowResult r = resolve_slow_warp(enc, sg, ws, arena, sb0, snb, arena_end, status, ids);   // same on all lanes
                if (lane == owner) mine = r;
            }
            return mine;
        }
    }
    const uint32_t src = __fns(m, 0, lane + 1);                                     // lane -> owner of the lane-th slow word
    const uint32_t sb0 = __shfl_sync(0xffffffffu, b0, src & 31), snb = __shfl_sync(0xffffffffu, nbytes, src & 31);
    SlowResult r; r.kind = kWordNone; r.ntok = 0; r.slot = 0; r.h6 = 0;
    if (lane < n_slow) r = resolve_slow(enc, sg, ws, arena, sb0, snb, arena_end, status);
    const uint32_t rank = __popc(m & ((1u << lane) - 1u));                          // this lane's word was resolved by lane `rank`
    mine.kind = __shfl_sync(0xffffffffu, r.kind, rank); mine.ntok = __shfl_sync(0xffffffffu, r.ntok, rank);
    mine.slot = __shfl_sync(0xffffffffu, r.slot, rank); mine.h6 = r.h6;
    return mine;
}

// pass 2 slow path: re-encode a short word whose ids were not kept
template <class Enc>
__device__ __noinline__ void emit_slow(const Enc &enc, const uint8_t *word, uint32_t nbytes, uint32_t kind, uint32_t ntok,
                                       uint32_t *dst) {
    (void)kind; (void)ntok;
    uint32_t buf[kShortBytes];
    uint32_t dummy = 0;
    const uint32_t n = enc.encode_short(nullptr, word, nbytes, buf, dummy);
    for (uint32_t k = 0; k < n; ++k) dst[k] = buf[k];
}

// ids of a one-load hit (16-bit ids, at most 13) -> dst.  The first eight are predicated stores without branches
// (a branch per token cost more than the stores); the rare second half sits behind one warp-uniform test.
template <bool kShared>
__device__ __forceinline__ void store_id_if(uint32_t *dst, uint32_t k, uint32_t n, uint32_t v) {
    if constexpr (kShared) {
        const uint32_t a = (uint32_t)__cvta_generic_to_shared(dst + k);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.gt.u32 p, %0, %1;\n\t@p st.shared.u32 [%2], %3;\n\t}" ::"r"(n), "r"(k), "r"(a), "r"(v) : "memory");
    } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.gt.u32 p, %0, %1;\n\t@p st.global.u32 [%2], %3;\n\t}" ::"r"(n), "r"(k), "l"(dst + k), "r"(v) : "memory");
    }
}
template <class Enc, bool kShared>
__device__ __forceinline__ void store_hit16_ids(uint32_t *dst, uint32_t n, uint4 a, uint4 b) {
    store_id_if<kShared>(dst, 0, n, Enc::expand16(a.x & 0xFFFFu, 0)); store_id_if<kShared>(dst, 1, n, Enc::expand16(a.x >> 16, 1));
    store_id_if<kShared>(dst, 2, n, Enc::expand16(a.y & 0xFFFFu, 2)); store_id_if<kShared>(dst, 3, n, Enc::expand16(a.y >> 16, 3));
    store_id_if<kShared>(dst, 4, n, Enc::expand16(a.z & 0xFFFFu, 4)); store_id_if<kShared>(dst, 5, n, Enc::expand16(a.z >> 16, 5));
    store_id_if<kShared>(dst, 6, n, Enc::expand16(a.w & 0xFFFFu, 6)); store_id_if<kShared>(dst, 7, n, Enc::expand16(a.w >> 16, 7));
    if (n > 8) {
        store_id_if<kShared>(dst, 8, n, Enc::expand16(b.x & 0xFFFFu, 8)); store_id_if<kShared>(dst, 9, n, Enc::expand16(b.x >> 16, 9));
        store_id_if<kShared>(dst, 10, n, Enc::expand16(b.y & 0xFFFFu, 10)); store_id_if<kShared>(dst, 11, n, Enc::expand16(b.y >> 16, 11));
        store_id_if<kShared>(dst, 12, n, Enc::expand16(b.z & 0xFFFFu, 12)); store_id_if<kShared>(dst, 13, n, Enc::expand16(b.z >> 16, 13));
        store_id_if<kShared>(dst, 14, n, Enc::expand16(b.w & 0xFFFFu, 14)); store_id_if<kShared>(dst, 15, n, Enc::expand16(b.w >> 16, 15));
    }
}

// clears the keys of the memo (ids16 / ext / tok32 are only read after a pub nibble was published) and the two cursors
static __global__ void __launch_bounds__(256) memo_clear_kernel(uint4 *keys, uint32_t n_slots, unsigned long long *long_cursor) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += gridDim.x * blockDim.x) keys[i] = z;
    if (blockIdx.x == 0 && threadIdx.x == 0) { long_cursor[0] = 0ull; long_cursor[1] = 0ull; }
}

// resolves `count` (<= 32) pending words of a warp, one per lane: record, token count into the tile total.  Returns the lane's
// H6 events.  Out of line: keeps the registers of the slow path out of the count loop.
template <class Enc>
__device__ __noinline__ uint32_t flush_pending_words(const Enc &enc, const typename Enc::Stage *sg, const EncodeWorkspace &ws, const uint8_t *arena,
                                                     const uint32_t *word_off, uint32_t arena_end, uint32_t *status, const uint32_t *pend,
                                                     uint32_t count, uint32_t warp_words, uint32_t *ids) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t h6 = 0;
    if constexpr (Enc::kWarpShort) {
        if (count <= warp_words) {                              // a few left-over words: the whole warp encodes them one after the other
            for (uint32_t k = 0; k < count; ++k) {
                const uint32_t w = pend[k];
                const uint32_t wb0 = __ldg(word_off + w), wnb = __ldg(word_off + w + 1) - wb0;
                const SlowResult r = resolve_slow_warp(enc, sg, ws, arena, wb0, wnb, arena_end, status, ids);
                if (lane == 0) {
                    ws.packed[w] = (r.kind << 29) | (r.ntok << 23) | ((r.kind == kWordHit || r.kind == kWordHit16) ? r.slot : 0u);
                    if (r.ntok) atomicAdd(&ws.tile_total[w / kTileWords], r.ntok);
                }
            }
            __syncwarp();
            return 0u;
        }
    }
    if (lane < count) {
        const uint32_t w = pend[lane];
        const uint32_t wb0 = __ldg(word_off + w), wnb = __ldg(word_off + w + 1) - wb0;
        const SlowResult r = resolve_slow(enc, sg, ws, arena, wb0, wnb, arena_end, status);
        ws.packed[w] = (r.kind << 29) | (r.ntok << 23) | ((r.kind == kWordHit || r.kind == kWordHit16) ? r.slot : 0u);
        if (r.ntok) atomicAdd(&ws.tile_total[w / kTileWords], r.ntok);
        h6 = r.h6;
    }
    __syncwarp();
    return h6;
}

// ---- pass 1: count ---------------------------------------------------------------------------------------------------
// Enc provides
//   struct Stage; void stage_init(Stage &) const   tables the CTA keeps in shared memory (FastBPE: Bloom filter + hot low-rank merges)
//   uint32_t encode_short(const Stage *, const uint8_t *p, uint32_t nbytes, uint32_t *buf /*thread-local, kShortBytes*/, uint32_t &h6) const
//   static bool narrow16(uint32_t id, uint32_t k, uint32_t &v16);  static uint32_t expand16(uint32_t v16, uint32_t k)
//   static constexpr bool kScratchLong, kBatchSlowPath, kWarpShort, kWarpLong
//   kWarpShort           : uint32_t encode_short_warp(const Stage *, p, nbytes, uint32_t *ids_smem /*32*/) const   (all 32 lanes, one word)
//   kScratchLong == false: uint32_t long_count(p, nbytes, h6) const;  void long_emit(p, nbytes, uint32_t *dst, uint32_t cap, uint32_t &h6) const
//   kWarpLong            : long_scratch_need(nbytes); long_count_warp(p, nbytes, scratch, h6); long_emit_warp(p, nbytes, scratch, dst, n)
//   kScratchLong == true : uint32_t encode_long_warp(p, nbytes, bufA, bufB, uint32_t **result) const   (all 32 lanes)
//
// Every warp owns tiles of kTileWords = 64 consecutive words (2 per lane), assigned round-robin; warps never wait for
// each other, so the L2 latencies of one warp's probes are covered by the other warps of the SM.
// Fast path of pass 1, one row = 32 consecutive words, word w_row + lane on lane `lane`.  The row's key bytes come from three aligned
// 8-byte loads per word (fewer L1 wavefronts than five 4-byte loads: this pass is bound by the L1 data pipe, not by the ALUs).
struct RowLoads { uint2 a, b, c; };
__device__ __forceinline__ RowLoads row_load(const uint8_t *arena_al, uint32_t pos /* byte offset of the word + phase of the arena */) {
    const uint2 *p = reinterpret_cast<const uint2 *>(arena_al + (pos & ~7u));
    RowLoads r;
    r.a = __ldg(p); r.b = __ldg(p + 1); r.c = __ldg(p + 2);
    return r;
}
// 128-bit memo key of a word of `nb` (1..32) bytes from the loaded words: first 15 bytes, zero padded, length nibble in the top byte
__device__ __forceinline__ uint4 row_key(const RowLoads &r, uint32_t pos, uint32_t nb) {
    const bool hi = (pos & 4u) != 0;
    const uint32_t sh = (pos & 3u) * 8u, n = min(nb, 15u);
    const uint32_t x0 = hi ? r.a.y : r.a.x, x1 = hi ? r.b.x : r.a.y, x2 = hi ? r.b.y : r.b.x, x3 = hi ? r.c.x : r.b.y, x4 = hi ? r.c.y : r.c.x;
    uint4 k;
    k.x = __funnelshift_r(x0, x1, sh) & low_bytes_mask((int)n);
    k.y = __funnelshift_r(x1, x2, sh) & low_bytes_mask((int)n - 4);
    k.z = __funnelshift_r(x2, x3, sh) & low_bytes_mask((int)n - 8);
    k.w = (__funnelshift_r(x3, x4, sh) & low_bytes_mask((int)n - 12)) | ((nb <= 15u ? nb : 0u) << 24);
    return k;
}
// (entry ^ key) without the pub nibble: 0 when the slot holds this key
__device__ __forceinline__ uint32_t key_diff(const uint4 &e, const uint4 &k) {
    return (e.x ^ k.x) | (e.y ^ k.y) | (e.z ^ k.z) | ((e.w ^ k.w) & 0x0FFFFFFFu);
}
template <class Enc>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
encode_count_kernel(Enc enc, const uint8_t *__restrict__ arena, const uint32_t *__restrict__ word_off, uint32_t n_words,
                    EncodeWorkspace ws, uint32_t *status, uint32_t warp_words, uint32_t tile_begin, uint32_t tile_end, uint32_t mode,
                    uint32_t warm_tiles) {
    // mode 0: plain.  1: the warm-up of the split count pass (its slow-path words are also counted in kStatusWarmSlow).  2: the rest of a
    // stream with many word types (runs only when the warm-up found one; otherwise the leaf / resolve kernels do the work).
    if (mode == 2 && !many_types_stream(status, warm_tiles)) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * kThreads + threadIdx.x) >> 5, n_warps = (gridDim.x * kThreads) >> 5;
    const uint32_t arena_end = word_off[n_words];
    const bool use_memo = ws.memo_mask != 0;
    const uint8_t *arena_al = reinterpret_cast<const uint8_t *>((uintptr_t)arena & ~(uintptr_t)7);   // 8-byte aligned base of the key loads
    const uint32_t aphase = (uint32_t)((uintptr_t)arena & 7);
    uint32_t h6 = 0;
    __shared__ typename Enc::Stage s_stage;
    enc.stage_init(s_stage);                                        // cooperative copy + __syncthreads (no-op for encoders without tables)
    const typename Enc::Stage *sg = &s_stage;

    // pending queue of the warp: words waiting for the slow path (at most 31 left over + 64 from one tile), followed by the 32-word
    // id staging row of the warp-per-word encoder (FastBPE)
    __shared__ uint32_t s_pend[kWarps][96 + 32];
    uint32_t *pend = s_pend[threadIdx.x >> 5];
    uint32_t n_pend = 0;
    uint32_t n_slow_words = 0;                                                      // diagnostic (status word 5), warp-uniform
    uint32_t n_gate_words = 0;                                                      // warm-up: slow-path words of its last quarter
    auto flush_pending = [&](uint32_t first, uint32_t count) {
        h6 += flush_pending_words(enc, sg, ws, arena, word_off, arena_end, status, pend + first, count, warp_words, pend + 96);
        n_slow_words += count;
    };
    // A tile is two rows of 32 words (word w_tile + 32 j + lane on lane `lane`): offsets, records and token offsets are coalesced 4-byte
    // accesses and the key loads of a row touch few lines.  The four offsets of the next tile are fetched one tile ahead.
    const uint32_t n_full = n_words / kTileWords;                                   // tiles below this one have all 64 words
    uint32_t po[4] = {0, 0, 0, 0};
    auto load_offsets = [&](uint32_t t) {
        const uint32_t *q = word_off + (size_t)t * kTileWords + lane;
        po[0] = __ldg(q); po[1] = __ldg(q + 1); po[2] = __ldg(q + 32); po[3] = __ldg(q + 33);
    };
    // The tile loop holds no calls: when 32 words are pending the warp leaves it, resolves them (flush_pending_words, out of line) and
    // enters it again, so that nothing of the loop's state (the prefetched offsets) has to survive a call.
    uint32_t tile = tile_begin + warp_global;
    while (tile < tile_end) {
    if (tile < n_full) load_offsets(tile);
    for (; tile < tile_end && n_pend < 32; tile += n_warps) {
        const uint32_t w_tile = tile * kTileWords;
        uint32_t rec[kWordsPerThread] = {0u, 0u}, ntok[kWordsPerThread] = {0u, 0u};
        bool slow[kWordsPerThread] = {false, false}, is_long[kWordsPerThread] = {false, false};
        const uint32_t o[4] = {po[0], po[1], po[2], po[3]};
        if (tile + n_warps < n_full) load_offsets(tile + n_warps);
        // the whole tile takes the fast path when it is full and 40 bytes can be read from the start of its last word
        const uint32_t last_start = __shfl_sync(0xffffffffu, o[2], 31);
        if (tile < n_full && use_memo && (uint64_t)last_start + 40 <= arena_end) {
            RowLoads ld[kWordsPerThread];
            uint32_t nb[kWordsPerThread], pos[kWordsPerThread], slot[kWordsPerThread];
            uint4 kw[kWordsPerThread], ew[kWordsPerThread];
#pragma unroll
            for (int j = 0; j < kWordsPerThread; ++j) { nb[j] = o[2 * j + 1] - o[2 * j]; pos[j] = o[2 * j] + aphase; ld[j] = row_load(arena_al, pos[j]); }
#pragma unroll
            for (int j = 0; j < kWordsPerThread; ++j) {
                kw[j] = row_key(ld[j], pos[j], nb[j]);
                slot[j] = memo_hash4(kw[j].x, kw[j].y, kw[j].z, kw[j].w) & ws.memo_mask;
                ew[j] = ld_ca_u32x4(ws.keys + slot[j]);              // ONE scattered load per word: key + pub nibble
            }
#pragma unroll
            for (int j = 0; j < kWordsPerThread; ++j) {
                // Served here: the first slot holds this word (at most 15 bytes, so the key is the whole word) with 1..13 tokens.
                // Everything else is rare and goes to the slow path, 32 words at a time: longer probe sequences, words of 16..32
                // bytes (tail compare), entries with 14..16 tokens or a 32-bit id list, first occurrences.  No calls in this loop.
                uint32_t d = key_diff(ew[j], kw[j]);
                if (d != 0 && (ew[j].x | ew[j].y | ew[j].z | ew[j].w) != 0) {       // another word lives here: try the neighbouring slot
                    slot[j] = (slot[j] + 1) & ws.memo_mask;
                    ew[j] = ld_ca_u32x4(ws.keys + slot[j]);
                    d = key_diff(ew[j], kw[j]);
                }
                const uint32_t pub = ew[j].w >> 28;
                const bool hit = d == 0 && pub - 1u < kPubWide - 1u && nb[j] - 1u < 15u;
                ntok[j] = hit ? pub : 0u;
                is_long[j] = nb[j] > (uint32_t)kShortBytes;
                slow[j] = !hit && !is_long[j];
                rec[j] = hit ? (kWordHit16 << 29) | (pub << 23) | slot[j] : is_long[j] ? kWordLong << 29 : 0u;
            }
        } else {
            // last tiles of the call (partial, or too close to the end of the arena for the unguarded loads) and calls without memo:
            // every word goes to the slow path, which builds its key bytewise
#pragma unroll
            for (int j = 0; j < kWordsPerThread; ++j) {
                const uint32_t w = w_tile + 32u * j + lane;
                if (w < n_words) {
                    const uint32_t nbj = __ldg(word_off + w + 1) - __ldg(word_off + w);
                    is_long[j] = nbj > (uint32_t)kShortBytes;
                    slow[j] = !is_long[j];
                    rec[j] = is_long[j] ? kWordLong << 29 : 0u;
                }
            }
        }
        // Words not served by the first probe (first occurrence, hash collision, ids that need the 32-bit list).  A single trie
        // walk / merge loop is a chain of dependent L2 accesses (10-25 us) during which the other lanes of the warp would idle.
        // kBatchSlowPath: they go to the warp's pending queue and are resolved 32 at a time, one per lane (flush below).
        // Otherwise (swt_tune("bpe_queue", 0)): a few slow words are encoded by the WHOLE WARP one after the other (kWarpShort: one
        // lane per adjacent pair, shuffle min-reduce); more than `warp_words` of them are spread over the lanes, one word per lane.
#pragma unroll
        for (int j = 0; j < kWordsPerThread; ++j) {
            const uint32_t m = __ballot_sync(0xffffffffu, slow[j]);
            if (m == 0) continue;                                                   // warp-uniform, the common case
            if (mode == 1 && tile >= tile_end - tile_end / 4u) n_gate_words += __popc(m);
            if constexpr (Enc::kBatchSlowPath) {
                if (slow[j]) pend[n_pend + __popc(m & ((1u << lane) - 1u))] = w_tile + 32u * j + lane;
                n_pend += __popc(m);
            } else {
                n_slow_words += __popc(m);
                uint32_t b0 = 0, nbj = 0;
                if (slow[j]) { b0 = __ldg(word_off + w_tile + 32u * j + lane); nbj = __ldg(word_off + w_tile + 32u * j + lane + 1) - b0; }
                const SlowResult r = resolve_slow_tile(enc, sg, ws, arena, arena_end, status, pend + 96, m, b0, nbj, warp_words);
                h6 += r.h6;
                if (slow[j]) {
                    ntok[j] = r.ntok; slow[j] = false;
                    rec[j] = (r.kind << 29) | (r.ntok << 23) | ((r.kind == kWordHit || r.kind == kWordHit16) ? r.slot : 0u);
                }
            }
        }
        // long words (rare): left to encode_long_count_kernel, which runs between this pass and the scan; here the tile is only
        // flagged, so that the code of the long paths (warp-cooperative merge loop / segment walk) stays out of this kernel
        if (__any_sync(0xffffffffu, is_long[0] || is_long[1])) { if (lane == 0) atomicOr(&ws.long_tiles[tile >> 5], 1u << (tile & 31u)); }
        // ---- per-word records and the tile total (pending words: record and token count are added by the flush)
#pragma unroll
        for (int j = 0; j < kWordsPerThread; ++j) {
            const uint32_t w = w_tile + 32u * j + lane;
            if (w < n_words && !slow[j]) ws.packed[w] = rec[j];
        }
        const uint32_t total = __reduce_add_sync(0xffffffffu, ntok[0] + ntok[1]);
        if (lane == 0) ws.tile_total[tile] = total;
    }
    __syncwarp();
    while (n_pend >= 32) { n_pend -= 32; flush_pending(n_pend, 32); }
    }
    if (n_pend) flush_pending(0, n_pend);
    if (h6) atomicAdd(&status[kStatusH6], h6);
    if (lane == 0 && n_slow_words) atomicAdd(&status[kStatusSlowWords], n_slow_words);
    if (lane == 0 && n_gate_words) atomicAdd(&status[kStatusWarmSlow], n_gate_words);
}

// ---- pass 1 without memo (swt_tune("memo_off", 1): the direct-path rates): every lane encodes its own words, no queue ----------------
template <class Enc>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
encode_count_direct_kernel(Enc enc, const uint8_t *__restrict__ arena, const uint32_t *__restrict__ word_off, uint32_t n_words, EncodeWorkspace ws,
                           uint32_t *status) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * kThreads + threadIdx.x) >> 5, n_warps = (gridDim.x * kThreads) >> 5;
    __shared__ typename Enc::Stage s_stage;
    enc.stage_init(s_stage);
    const typename Enc::Stage *sg = &s_stage;
    uint32_t h6 = 0;
    for (uint32_t tile = warp_global; tile < ws.n_tiles; tile += n_warps) {
        uint32_t total = 0; bool any_long = false;
#pragma unroll
        for (int j = 0; j < kWordsPerThread; ++j) {
            const uint32_t w = tile * kTileWords + 32u * j + lane;
            uint32_t ntok = 0;
            if (w < n_words) {
                const uint32_t b0 = __ldg(word_off + w), nb = __ldg(word_off + w + 1) - b0;
                if (nb > (uint32_t)kShortBytes) { any_long = true; ws.packed[w] = kWordLong << 29; }
                else {
                    uint32_t buf[kShortBytes];
                    ntok = enc.encode_short(sg, arena + b0, nb, buf, h6);
                    ws.packed[w] = (kWordRecompute << 29) | (ntok << 23);
                }
            }
            total += ntok;
        }
        if (__any_sync(0xffffffffu, any_long)) { if (lane == 0) atomicOr(&ws.long_tiles[tile >> 5], 1u << (tile & 31u)); }
        total = __reduce_add_sync(0xffffffffu, total);
        if (lane == 0) ws.tile_total[tile] = total;
    }
    if (h6) atomicAdd(&status[kStatusH6], h6);
    if (lane == 0 && warp_global == 0) atomicAdd(&status[kStatusSlowWords], n_words);
}

// ---- pass 1, split form: leaf count kernel + resolve kernel ------------------------------------------------------------------------
// The kernel above carries the slow path (calls, a stack frame, the encoder's tables): the FastBPE instantiations want 98 registers,
// get 64 and spill inside the tile loop (count pass 1.40 ms per GB against 0.90 ms for FastWP, same fast path).  Once the memo is warm
// almost every word is served by the first probe, so for the encoders with Enc::kSplitCount the bulk of the stream goes through a
// LEAF kernel that only probes -- no encoder, no calls, no stack, 39 registers, 6 CTAs per SM -- and leaves the words it cannot serve
// as one bit per word in tile_pend[]; encode_resolve_kernel then walks those bits and resolves the words 32 at a time with the same
// flush_pending_words as above.  The stream is cut into chunks (count, resolve, count, resolve) so that a type first seen in one chunk
// is published before the next chunk is counted.  Measured per GB: FastBPE 1.40 -> 1.15 (1 chunk) / 1.21 ms (2 chunks); FastWP
// 0.90 -> 0.94 / 0.97 ms (the resolve kernels are exposed latency, which the single kernel hides behind other warps' tiles), so
// FastWP keeps the single kernel.
constexpr int kLeafCtasPerSm = 6;
static __global__ void __launch_bounds__(kThreads, kLeafCtasPerSm)
encode_count_leaf_kernel(const uint8_t *__restrict__ arena, const uint32_t *__restrict__ word_off, uint32_t n_words, EncodeWorkspace ws,
                         const uint32_t *status, uint32_t tile_begin, uint32_t tile_end, uint32_t warm_tiles) {
    if (many_types_stream(status, warm_tiles)) return;              // such streams stay with the kernel that resolves its misses itself
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * kThreads + threadIdx.x) >> 5, n_warps = (gridDim.x * kThreads) >> 5;
    const uint32_t arena_end = word_off[n_words];
    const uint8_t *arena_al = reinterpret_cast<const uint8_t *>((uintptr_t)arena & ~(uintptr_t)7);
    const uint32_t aphase = (uint32_t)((uintptr_t)arena & 7);
    const uint32_t n_full = n_words / kTileWords;
    uint32_t po[4] = {0, 0, 0, 0};
    auto load_offsets = [&](uint32_t t) {
        const uint32_t *q = word_off + (size_t)t * kTileWords + lane;
        po[0] = __ldg(q); po[1] = __ldg(q + 1); po[2] = __ldg(q + 32); po[3] = __ldg(q + 33);
    };
    uint32_t tile = tile_begin + warp_global;
    if (tile < tile_end && tile < n_full) load_offsets(tile);
    for (; tile < tile_end; tile += n_warps) {
        const uint32_t w_tile = tile * kTileWords;
        uint32_t rec[kWordsPerThread] = {0u, 0u}, ntok[kWordsPerThread] = {0u, 0u};
        bool slow[kWordsPerThread] = {false, false}, is_long[kWordsPerThread] = {false, false};
        const uint32_t o[4] = {po[0], po[1], po[2], po[3]};
        if (tile + n_warps < tile_end && tile + n_warps < n_full) load_offsets(tile + n_warps);
        const uint32_t last_start = __shfl_sync(0xffffffffu, o[2], 31);
        if (tile < n_full && (uint64_t)last_start + 40 <= arena_end) {
            RowLoads ld[kWordsPerThread];
            uint32_t nb[kWordsPerThread], pos[kWordsPerThread], slot[kWordsPerThread];
            uint4 kw[kWordsPerThread], ew[kWordsPerThread];
#pragma unroll
            for (int j = 0; j < kWordsPerThread; ++j) { nb[j] = o[2 * j + 1] - o[2 * j]; pos[j] = o[2 * j] + aphase; ld[j] = row_load(arena_al, pos[j]); }
#pragma unroll
            for (int j = 0; j < kWordsPerThread; ++j) {
                kw[j] = row_key(ld[j], pos[j], nb[j]);
                slot[j] = memo_hash4(kw[j].x, kw[j].y, kw[j].z, kw[j].w) & ws.memo_mask;
                ew[j] = ld_ca_u32x4(ws.keys + slot[j]);
            }
#pragma unroll
            for (int j = 0; j < kWordsPerThread; ++j) {
                uint32_t d = key_diff(ew[j], kw[j]);
                if (d != 0 && (ew[j].x | ew[j].y | ew[j].z | ew[j].w) != 0) {       // another word lives here: try the neighbouring slot
                    slot[j] = (slot[j] + 1) & ws.memo_mask;
                    ew[j] = ld_ca_u32x4(ws.keys + slot[j]);
                    d = key_diff(ew[j], kw[j]);
                }
                const uint32_t pub = ew[j].w >> 28;
                const bool hit = d == 0 && pub - 1u < kPubWide - 1u && nb[j] - 1u < 15u;
                ntok[j] = hit ? pub : 0u;
                is_long[j] = nb[j] > (uint32_t)kShortBytes;
                slow[j] = !hit && !is_long[j];
                rec[j] = hit ? (kWordHit16 << 29) | (pub << 23) | slot[j] : is_long[j] ? kWordLong << 29 : 0u;
            }
        } else {
#pragma unroll
            for (int j = 0; j < kWordsPerThread; ++j) {
                const uint32_t w = w_tile + 32u * j + lane;
                if (w < n_words) {
                    const uint32_t nbj = __ldg(word_off + w + 1) - __ldg(word_off + w);
                    is_long[j] = nbj > (uint32_t)kShortBytes;
                    slow[j] = !is_long[j];
                    rec[j] = is_long[j] ? kWordLong << 29 : 0u;
                }
            }
        }
        const uint32_t m0 = __ballot_sync(0xffffffffu, slow[0]), m1 = __ballot_sync(0xffffffffu, slow[1]);
        if ((m0 | m1) != 0 && lane == 0) ws.tile_pend[tile] = make_uint2(m0, m1);   // (the array was zeroed by the launcher)
        if (__any_sync(0xffffffffu, is_long[0] || is_long[1])) { if (lane == 0) atomicOr(&ws.long_tiles[tile >> 5], 1u << (tile & 31u)); }
#pragma unroll
        for (int j = 0; j < kWordsPerThread; ++j) {
            const uint32_t w = w_tile + 32u * j + lane;
            if (w < n_words && !slow[j]) ws.packed[w] = rec[j];
        }
        const uint32_t total = __reduce_add_sync(0xffffffffu, ntok[0] + ntok[1]);
        if (lane == 0) ws.tile_total[tile] = total;
    }
}

// resolves the words the leaf kernel left behind: a warp reads the pending masks of 32 tiles at a time, queues the words and flushes
// the queue 32 words at a time (records and token counts are written by the flush)
template <class Enc>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
encode_resolve_kernel(Enc enc, const uint8_t *__restrict__ arena, const uint32_t *__restrict__ word_off, uint32_t n_words, EncodeWorkspace ws,
                      uint32_t *status, uint32_t warp_words, uint32_t tile_begin, uint32_t tile_end, uint32_t warm_tiles) {
    if (many_types_stream(status, warm_tiles)) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * kThreads + threadIdx.x) >> 5, n_warps = (gridDim.x * kThreads) >> 5;
    const uint32_t arena_end = word_off[n_words];
    __shared__ typename Enc::Stage s_stage;
    enc.stage_init(s_stage);
    const typename Enc::Stage *sg = &s_stage;
    __shared__ uint32_t s_pend[kWarps][96 + 32];
    uint32_t *pend = s_pend[threadIdx.x >> 5];
    uint32_t n_pend = 0, n_slow_words = 0, h6 = 0;
    for (uint32_t tb = tile_begin + warp_global * 32u; tb < tile_end; tb += n_warps * 32u) {
        const uint32_t my_tile = tb + lane;
        const uint2 pm = my_tile < tile_end ? ws.tile_pend[my_tile] : make_uint2(0u, 0u);
        for (uint32_t nz = __ballot_sync(0xffffffffu, (pm.x | pm.y) != 0); nz; nz &= nz - 1) {
            const uint32_t src = __ffs(nz) - 1, w_tile = (tb + src) * kTileWords;
#pragma unroll
            for (int j = 0; j < kWordsPerThread; ++j) {
                const uint32_t m = __shfl_sync(0xffffffffu, j == 0 ? pm.x : pm.y, src);
                if (m == 0) continue;                                               // warp-uniform
                if ((m >> lane) & 1u) pend[n_pend + __popc(m & ((1u << lane) - 1u))] = w_tile + 32u * j + lane;
                n_pend += __popc(m);
                __syncwarp();
                while (n_pend >= 32) {
                    n_pend -= 32;
                    h6 += flush_pending_words(enc, sg, ws, arena, word_off, arena_end, status, pend + n_pend, 32, warp_words, pend + 96);
                    n_slow_words += 32;
                }
            }
        }
    }
    if (n_pend) { h6 += flush_pending_words(enc, sg, ws, arena, word_off, arena_end, status, pend, n_pend, warp_words, pend + 96); n_slow_words += n_pend; }
    if (h6) atomicAdd(&status[kStatusH6], h6);
    if (lane == 0 && n_slow_words) atomicAdd(&status[kStatusSlowWords], n_slow_words);
}

// ---- long words (> kShortBytes), between pass 1 and the scan -------------------------------------------------------------------
// One warp per flagged tile: for every long word of the tile the whole warp runs the encoder's long path, writes the word's final
// record and adds its tokens to the tile total.
//   kScratchLong (BPE): merge loop over symbols in global scratch, result kept there (record = scratch granule)
//   kWarpLong (FastWP) with scratch: the chunk's segments are walked by different lanes (record = granule of the segment records)
//   otherwise: the owning lane walks the word alone (record = token count; walked again by encode_long_emit_kernel)
template <class Enc>
__global__ void __launch_bounds__(256) encode_long_count_kernel(Enc enc, const uint8_t *__restrict__ arena, const uint32_t *__restrict__ word_off,
                                                                uint32_t n_words, EncodeWorkspace ws, uint32_t *status) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n_bm = (ws.n_tiles + 31) >> 5;
    uint32_t h6 = 0;
    // the bitmap is read 32 words per warp step (one per lane); nearly all of it is zero
    for (uint32_t bw0 = warp_global * 32u; bw0 < n_bm; bw0 += n_warps * 32u) {
      const uint32_t my_bits = bw0 + lane < n_bm ? ws.long_tiles[bw0 + lane] : 0u;
      for (uint32_t nz = __ballot_sync(0xffffffffu, my_bits != 0); nz; nz &= nz - 1) {
        const uint32_t src = __ffs(nz) - 1, bw = bw0 + src;
        uint32_t bits = __shfl_sync(0xffffffffu, my_bits, src);                     // warp-uniform
        while (bits) {
            const uint32_t tile = (bw << 5) + (__ffs(bits) - 1); bits &= bits - 1;
            const uint32_t w_tile = tile * kTileWords, tile_words = min((uint32_t)kTileWords, n_words - w_tile);
            uint32_t added = 0;
            for (uint32_t k = 0; k < tile_words; ++k) {                             // warp-uniform loop over the tile's words
                const uint32_t w = w_tile + k;
                const uint32_t lb0 = __ldg(word_off + w), lnb = __ldg(word_off + w + 1) - lb0;
                if (lnb <= (uint32_t)kShortBytes) continue;
                uint32_t rec = 0, c = 0;
                if constexpr (Enc::kScratchLong) {
                    const unsigned long long need = (kLongHeader + 2ull * lnb + 15ull) & ~15ull;
                    unsigned long long so = 0;
                    if (lane == 0) so = atomicAdd(ws.long_cursor, need);
                    so = __shfl_sync(0xffffffffu, so, 0);
                    const bool fits = so + need <= ws.long_scratch_elems && (so >> 4) < (1ull << 29);
                    uint32_t *res = nullptr;
                    uint32_t *bufA = ws.long_scratch + so + kLongHeader;
                    if (fits) {
                        c = enc.encode_long_warp(arena + lb0, lnb, bufA, bufA + lnb, &res);
                        if (lane == 0) ws.long_scratch[so] = c;
                        rec = ((res == bufA ? kWordLong : kWordLongB) << 29) | (uint32_t)(so >> 4);
                    } else if (lane == 0) atomicExch(&status[kStatusCode], (uint32_t)SWT_ERR_CAPACITY);
                } else {
                    bool done = false;
                    if constexpr (Enc::kWarpLong) {
                        const unsigned long long need = Enc::long_scratch_need(lnb);
                        unsigned long long so = 0;
                        if (lane == 0 && ws.long_scratch_elems) so = atomicAdd(ws.long_cursor, need);
                        so = __shfl_sync(0xffffffffu, so, 0);
                        if (ws.long_scratch_elems && so + need <= ws.long_scratch_elems && (so >> 4) < (1ull << 29)) {
                            uint32_t lh6 = 0;
                            c = enc.long_count_warp(arena + lb0, lnb, ws.long_scratch + so, lh6);
                            h6 += lh6;
                            rec = (kWordLongSeg << 29) | (uint32_t)(so >> 4);
                            done = true;
                        }
                    }
                    if (!done) {
                        if (lane == 0) c = enc.long_count(arena + lb0, lnb, h6);
                        c = __shfl_sync(0xffffffffu, c, 0);
                        if (c >= (1u << 29)) { if (lane == 0) atomicExch(&status[kStatusCode], (uint32_t)SWT_ERR_CAPACITY); c = 0; }
                        rec = (kWordLong << 29) | c;
                    }
                }
                if (lane == 0) ws.packed[w] = rec;
                added += c;
            }
            if (lane == 0 && added) atomicAdd(&ws.tile_total[tile], added);
        }
      }
    }
    if (h6) atomicAdd(&status[kStatusH6], h6);
}

// token count of a per-word record (after encode_long_count_kernel)
template <class Enc>
__device__ __forceinline__ uint32_t record_ntok(uint32_t packed, const uint32_t *long_scratch) {
    const uint32_t kind = packed >> 29, arg = packed & 0x1FFFFFFFu;
    if (kind == kWordHit || kind == kWordHit16 || kind == kWordRecompute) return arg >> 23;
    if (kind == kWordLongSeg || (Enc::kScratchLong && (kind == kWordLong || kind == kWordLongB))) return long_scratch[(unsigned long long)arg << 4];
    if (kind == kWordLong) return arg;
    return 0u;
}

// ---- long words, after pass 2: their ids are written at their final position (pass 2 left the gap) ------------------------------
template <class Enc>
__global__ void __launch_bounds__(256) encode_long_emit_kernel(Enc enc, const uint8_t *__restrict__ arena, const uint32_t *__restrict__ word_off,
                                                               uint32_t n_words, uint32_t *__restrict__ out_ids, uint64_t out_cap, EncodeWorkspace ws) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n_bm = (ws.n_tiles + 31) >> 5;
    for (uint32_t bw0 = warp_global * 32u; bw0 < n_bm; bw0 += n_warps * 32u) {
      const uint32_t my_bits = bw0 + lane < n_bm ? ws.long_tiles[bw0 + lane] : 0u;
      for (uint32_t nz = __ballot_sync(0xffffffffu, my_bits != 0); nz; nz &= nz - 1) {
        const uint32_t src = __ffs(nz) - 1, bw = bw0 + src;
        uint32_t bits = __shfl_sync(0xffffffffu, my_bits, src);
        while (bits) {
            const uint32_t tile = (bw << 5) + (__ffs(bits) - 1); bits &= bits - 1;
            const uint32_t w_tile = tile * kTileWords, tile_words = min((uint32_t)kTileWords, n_words - w_tile);
            const uint64_t base = ws.group_base[tile / kGroupTiles] + ws.tile_total[tile];
            // tile-local token offsets of the words: lane holds words 2*lane, 2*lane + 1 like pass 2
            const uint32_t i0 = lane * kWordsPerThread;
            uint32_t pk[kWordsPerThread], nt[kWordsPerThread], run[kWordsPerThread], count = 0;
#pragma unroll
            for (int j = 0; j < kWordsPerThread; ++j) {
                pk[j] = i0 + j < tile_words ? ws.packed[w_tile + i0 + j] : 0u;
                nt[j] = record_ntok<Enc>(pk[j], ws.long_scratch);
                run[j] = count; count += nt[j];
            }
            uint32_t incl = count;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += v; }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            if (base + total > out_cap) continue;                                   // status set by the scan
#pragma unroll
            for (int j = 0; j < kWordsPerThread; ++j) {
                run[j] += incl - count;
                const uint32_t kj = pk[j] >> 29;
                uint32_t lm = __ballot_sync(0xffffffffu, kj == kWordLong || kj == kWordLongB || kj == kWordLongSeg);
                while (lm) {
                    const uint32_t owner = __ffs(lm) - 1; lm &= lm - 1;
                    const uint32_t ln = __shfl_sync(0xffffffffu, nt[j], owner), lrun = __shfl_sync(0xffffffffu, run[j], owner);
                    const uint32_t lpk = __shfl_sync(0xffffffffu, pk[j], owner);
                    const uint32_t lkind = lpk >> 29, larg = lpk & 0x1FFFFFFFu;
                    const uint32_t w = w_tile + owner * kWordsPerThread + j;
                    const uint32_t lb0 = __ldg(word_off + w), lnb = __ldg(word_off + w + 1) - lb0;
                    uint32_t *dst = out_ids + base + lrun;
                    if constexpr (Enc::kScratchLong) {
                        const uint32_t *src = ws.long_scratch + ((unsigned long long)larg << 4) + kLongHeader + (lkind == kWordLongB ? lnb : 0u);
                        for (uint32_t k = lane; k < ln; k += 32) dst[k] = src[k];
                    } else {
                        bool done = false;
                        if constexpr (Enc::kWarpLong) {
                            if (lkind == kWordLongSeg) { enc.long_emit_warp(arena + lb0, lnb, ws.long_scratch + ((unsigned long long)larg << 4), dst, ln); done = true; }
                        }
                        if (!done && lane == 0) { uint32_t dummy = 0; enc.long_emit(arena + lb0, lnb, dst, ln, dummy); }
                    }
                }
            }
        }
      }
    }
}

// ---- scan of the tile totals: in-group exclusive prefixes (in place) + group sums, then the group bases ------------------
static __global__ void __launch_bounds__(256) encode_scan_groups_kernel(EncodeWorkspace ws) {
    __shared__ uint32_t sh_scan[36];
    const uint32_t g = blockIdx.x, t0 = g * kGroupTiles + threadIdx.x * 4;
    uint32_t v[4], s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = t0 + k < ws.n_tiles ? ws.tile_total[t0 + k] : 0u; s += v[k]; }
    uint32_t total, excl = block_exclusive_scan(s, sh_scan, &total);
#pragma unroll
    for (int k = 0; k < 4; ++k) { if (t0 + k < ws.n_tiles) ws.tile_total[t0 + k] = excl; excl += v[k]; }
    if (threadIdx.x == 0) ws.group_base[g] = total;
}
static __global__ void __launch_bounds__(1024) encode_scan_top_kernel(EncodeWorkspace ws, uint32_t n_words, uint32_t *out_tok_off, uint32_t tok_base,
                                                               uint64_t out_cap, uint32_t *status) {
    __shared__ unsigned long long sh[33];
    __shared__ unsigned long long carry;
    const uint32_t n_groups = (ws.n_tiles + kGroupTiles - 1) / kGroupTiles;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t g0 = 0; g0 < n_groups; g0 += 1024) {
        const uint32_t g = g0 + threadIdx.x;
        const unsigned long long v = g < n_groups ? ws.group_base[g] : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned long long u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += u; }
        if (lane == 31) sh[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            unsigned long long w = sh[lane], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const unsigned long long u = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= (uint32_t)d) wi += u; }
            sh[lane] = wi - w;
            if (lane == 31) sh[32] = wi;
        }
        __syncthreads();
        const unsigned long long c = carry;
        if (g < n_groups) ws.group_base[g] = c + sh[wid] + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + sh[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const unsigned long long grand = carry;
        if (out_tok_off) out_tok_off[n_words] = tok_base + (uint32_t)grand;
        status[kStatusTokens] = (uint32_t)grand; status[kStatusTokensHi] = (uint32_t)(grand >> 32);
        if (grand > out_cap) atomicExch(&status[kStatusCode], (uint32_t)SWT_ERR_CAPACITY);
    }
}

// ---- pass 2: emit ------------------------------------------------------------------------------------------------------
// Bulk copy-out: the 16-byte aligned middle of a tile's ids goes shared -> global as one cp.async.bulk issued by
// lane 0 (TMA engine; no LDS/STG wavefronts on the LSU pipe that bounds this kernel), the at most three ids in front of / behind it
// as scalar stores.  The buffer is reused by the next tile after cp.async.bulk.wait_group.read.
__device__ __forceinline__ void bulk_store_s2g(void *gdst, const void *ssrc, uint32_t bytes) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(sa), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Tiles of pass 2 with an unusual record (32-bit id list of the tok32 arena, a word that has to be encoded again, a long word whose
// ids encode_long_emit_kernel writes) or with more tokens than the shared-memory buffer holds: rare, out of line, ids written straight
// to global memory.  ntok[] / run[] are recomputed here because the counts of long words do not fit the packed scan of the fast path.
template <class Enc>
__device__ __noinline__ void emit_tile_generic(const Enc &enc, const uint8_t *arena, const uint32_t *word_off, uint32_t n_words, uint32_t *out_ids,
                                               uint64_t out_cap, uint32_t *out_tok_off, uint32_t tok_base, const EncodeWorkspace &ws, uint32_t w_tile,
                                               uint64_t base, uint32_t p0, uint32_t p1) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t packed[kWordsPerThread] = {p0, p1};
    uint32_t ntok[kWordsPerThread], run[kWordsPerThread], total = 0;
#pragma unroll
    for (int j = 0; j < kWordsPerThread; ++j) {
        ntok[j] = record_ntok<Enc>(packed[j], ws.long_scratch);
        uint32_t incl = ntok[j];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += v; }
        run[j] = total + incl - ntok[j];
        total += __shfl_sync(0xffffffffu, incl, 31);
    }
#pragma unroll
    for (int j = 0; j < kWordsPerThread; ++j) {
        const uint32_t w = w_tile + 32u * j + lane;
        if (out_tok_off && w < n_words) out_tok_off[w] = tok_base + (uint32_t)base + run[j];
    }
    if (base + total > out_cap) return;                                             // status set by the scan
#pragma unroll
    for (int j = 0; j < kWordsPerThread; ++j) {
        const uint32_t kind = packed[j] >> 29, arg = packed[j] & 0x1FFFFFFFu, w = w_tile + 32u * j + lane;
        uint32_t *dst = out_ids + base + run[j];
        if (kind == kWordHit16) {
            const uint4 *e = ws.ids16 + 2 * (size_t)(arg & kMemoSlotMask);
            const uint4 a = ld_ca_u32x4(e), b = ntok[j] > 8 ? ld_ca_u32x4(e + 1) : make_uint4(0, 0, 0, 0);
            store_hit16_ids<Enc, false>(dst, ntok[j], a, b);
        } else if (kind == kWordHit) {                                              // 32-bit id list in the tok32 arena
            const MemoExt *x = ws.ext + (arg & kMemoSlotMask);
            if (x->meta != 0xFFFFFFFFu) { const uint32_t *src = ws.tok32 + x->tok32_off; for (uint32_t k = 0; k < ntok[j]; ++k) dst[k] = src[k]; }
            else {                                                                  // the owner found the arena full: encode again
                const uint32_t b0 = __ldg(word_off + w), b1 = __ldg(word_off + w + 1);
                emit_slow(enc, arena + b0, b1 - b0, kind, ntok[j], dst);
            }
        } else if (kind == kWordRecompute) {
            const uint32_t b0 = __ldg(word_off + w), b1 = __ldg(word_off + w + 1);
            emit_slow(enc, arena + b0, b1 - b0, kind, ntok[j], dst);
        }
    }
}

template <class Enc>
__global__ void __launch_bounds__(kThreads, kEmitCtasPerSm)
encode_emit_kernel(Enc enc, const uint8_t *__restrict__ arena, const uint32_t *__restrict__ word_off, uint32_t n_words,
                   uint32_t *__restrict__ out_ids, uint64_t out_cap, uint32_t *__restrict__ out_tok_off, uint32_t tok_base,
                   EncodeWorkspace ws) {
    __shared__ __align__(16) uint32_t s_compact[kWarps][kCompactTokens + 8];   // per warp: the tile's ids in output order
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * kThreads + threadIdx.x) >> 5, n_warps = (gridDim.x * kThreads) >> 5;
    uint32_t *compact = s_compact[threadIdx.x >> 5];
    bool bulk_pending = false;                                                  // lane 0: a bulk copy may still be reading `compact`

    // software pipeline: the per-word records (two rows of 32 words, word w_tile + 32 j + lane on lane `lane`) and the tile's output
    // position are fetched two tiles ahead (the two addends of the position are kept apart: adding them at load time made every
    // iteration wait for the loads it had just issued)
    auto load_tile = [&](uint32_t t, uint32_t &p0, uint32_t &p1, unsigned long long &gb, uint32_t &tt) {
        const uint32_t w = t * kTileWords + lane;
        p0 = w < n_words ? ws.packed[w] : 0u;
        p1 = w + 32 < n_words ? ws.packed[w + 32] : 0u;
        gb = ws.group_base[t / kGroupTiles]; tt = ws.tile_total[t];
    };
    uint32_t pp0 = 0, pp1 = 0, ptt = 0, qp0 = 0, qp1 = 0, qtt = 0; unsigned long long pgb = 0, qgb = 0;
    if (warp_global < ws.n_tiles) load_tile(warp_global, pp0, pp1, pgb, ptt);
    if (warp_global + n_warps < ws.n_tiles) load_tile(warp_global + n_warps, qp0, qp1, qgb, qtt);
    for (uint32_t tile = warp_global; tile < ws.n_tiles; tile += n_warps) {
        const uint32_t w_tile = tile * kTileWords;
        const uint32_t p0 = pp0, p1 = pp1;
        const uint64_t base = pgb + ptt;
        pp0 = qp0; pp1 = qp1; pgb = qgb; ptt = qtt;
        if (tile + 2 * n_warps < ws.n_tiles) load_tile(tile + 2 * n_warps, qp0, qp1, qgb, qtt);
        // records are kWordNone (no word / no token) or kWordHit16 in all but a few tiles
        if (__any_sync(0xffffffffu, (p0 | p1) >= (kWordHit << 29))) {
            emit_tile_generic(enc, arena, word_off, n_words, out_ids, out_cap, out_tok_off, tok_base, ws, w_tile, base, p0, p1);
            continue;
        }
        // ids of both rows: one scattered 16-byte load per word (a second one for words of more than 8 tokens), all in flight together
        const uint32_t n0 = (p0 >> 23) & 63u, n1 = (p1 >> 23) & 63u;                // kWordNone records are zero
        const uint4 *e0 = ws.ids16 + 2 * (size_t)(p0 & kMemoSlotMask), *e1 = ws.ids16 + 2 * (size_t)(p1 & kMemoSlotMask);
        uint4 a0 = make_uint4(0, 0, 0, 0), b0 = a0, a1 = a0, b1 = a0;
        if (n0) a0 = ld_ca_u32x4(e0);
        if (n1) a1 = ld_ca_u32x4(e1);
        if (n0 > 8) b0 = ld_ca_u32x4(e0 + 1);
        if (n1 > 8) b1 = ld_ca_u32x4(e1 + 1);
        // tile-local token offsets: ONE inclusive scan over both rows' counts packed into 16-bit halves (a row has at most 32 * 16 tokens)
        uint32_t incl = n0 | (n1 << 16);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += v; }
        const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t tot0 = tot & 0xFFFFu, total = tot0 + (tot >> 16);
        const uint32_t run0 = (incl & 0xFFFFu) - n0, run1 = tot0 + (incl >> 16) - n1;     // row 0 comes first in the output
        if (out_tok_off) {
            const uint32_t o0 = tok_base + (uint32_t)base, w = w_tile + lane;
            if (w < n_words) out_tok_off[w] = o0 + run0;
            if (w + 32 < n_words) out_tok_off[w + 32] = o0 + run1;
        }
        if (base + total > out_cap) continue;                                       // warp-uniform (status set by the scan)
        if (total > (uint32_t)kCompactTokens) {                                     // more tokens than the buffer holds (rare)
            emit_tile_generic(enc, arena, word_off, n_words, out_ids, out_cap, nullptr, tok_base, ws, w_tile, base, p0, p1);
            continue;
        }
        // the previous tile's bulk copy must have finished READING the buffer before it is overwritten
        if (lane == 0 && bulk_pending) { bulk_store_wait_read(); bulk_pending = false; }
        __syncwarp();
        // The tile's ids are laid out in shared memory with the 16-byte phase of their destination and leave as ONE bulk copy
        // (cp.async.bulk shared::cta -> global: no LDS/STG wavefronts on the LSU pipe that bounds this kernel); the at most three ids in
        // front of / behind the 16-byte aligned middle are plain stores of lanes 1..6.
        const uint32_t sh = (uint32_t)((uintptr_t)(out_ids + base) >> 2) & 3u;
        uint32_t *cdst = compact + sh;
        store_hit16_ids<Enc, true>(cdst + run0, n0, a0, b0);
        store_hit16_ids<Enc, true>(cdst + run1, n1, a1, b1);
        fence_proxy_async_smem();                                                   // generic-proxy writes -> visible to the async proxy
        __syncwarp();
        uint32_t *dsta = out_ids + base - sh;                                       // 16-byte aligned
        const uint32_t end = sh + total;
        const uint32_t c0 = sh ? 4u : 0u, c1 = end & ~3u;                           // aligned middle [c0, c1)
        if (lane == 0 && c1 > c0) { bulk_store_s2g(dsta + c0, compact + c0, (c1 - c0) * 4u); bulk_pending = true; }
        {   // head [sh, min(4, end)) on lanes 1..3, tail [max(c1, c0), end) on lanes 4..6
            const uint32_t c = lane < 4 ? lane : c1 + lane - 4;
            const bool head = lane < 4 && c >= sh && sh != 0, tail = lane >= 4 && lane < 7 && c1 >= c0;
            if ((head || tail) && c < end) dsta[c] = compact[c];
        }
        __syncwarp();                                            // the compact buffer is reused by the next tile
    }
    if (lane == 0 && bulk_pending) bulk_store_wait_read();       // shared memory must stay valid until the last copy has read it
}

// ---- small calls: ONE CTA pre-tokenizes and encodes one short text (swt_tokenize_small) -------------------------------------------
// The call pattern of the reference's CLI is one tokenize() per line (cli.py:253-264): a few dozen bytes per call.  The batch kernels
// above cost a dozen launches, a memo clear and three synchronisations; here a single CTA does everything in one launch:
//   1. pre-tokenizer (pretok.cuh): one warp per 4 KiB tile counts, thread 0 scans the tile sums, the warps write the lower-cased word
//      arena + offsets to device scratch.  The text is read from mapped pinned host memory (zero copy).
//   2. encoder: every word directly, one thread per word, no memo (rank table / trie in L2).  Texts of up to 256 words without a
//      word longer than kShortBytes keep the ids in the thread and write them straight to `out`; otherwise the ids go through
//      scratch (ids of word w at scratch[off[w] + w ...]: a word has at most max(bytes, 1) tokens) and a compacted copy.
//   out (mapped pinned host memory): [0] status code, [1] tokens, [2] H6 events, [3] words, [4] sequence number of the call (the
//   completion flag), [5..7] SM cycles from kernel start to the end of the pre-tokenizer / the encoder / the id stores, ids from out[8]
struct SmallArgs {
    uint8_t *arena; uint32_t *word_off;            // device scratch: lower-cased words of the text, n_words + 1 offsets
    uint32_t *scratch, *cnt, *compact, *long_buf;  // u32[arena + words + 1], u32[words], u32[tokens + 4], u32[2 * arena + 32] (BPE long words)
    uint32_t *out; uint32_t out_cap;
    uint32_t seq;                                  // written to out[4] LAST (system-scope fence before it): the host polls this word
};
constexpr int kSmallThreads = 256;
// texts of up to kInlineTextBytes travel in the kernel's parameter buffer (no read over PCIe at all)
constexpr uint32_t kInlineTextBytes = 240;
struct InlineText { uint4 v[kInlineTextBytes / 16 + 1]; };
template <class Enc, bool kBert, bool kInline>
__global__ void __launch_bounds__(kSmallThreads) tokenize_small_kernel(Enc enc, pt::PretokDev t, const uint8_t *__restrict__ text, InlineText inl, uint32_t n,
                                                                       SmallArgs a) {
    __shared__ unsigned long long s_sum[pt::kSmallTiles + 1];
    __shared__ uint32_t s_status[8], sh_scan[36], s_h6, s_carry;
    __shared__ __align__(16) uint8_t s_text[pt::kSmallTiles * pt::kTileBytes + 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long clk0 = clock64();                               // diagnostic: SM cycles of the phases go to out[5..7]
    // ---- 1. pre-tokenizer
    const uint32_t n_tiles = (n + pt::kTileBytes - 1) / pt::kTileBytes;
    if (tid < 8) s_status[tid] = 0u;
    if (tid == 0) { s_h6 = 0; s_carry = 0; }
    __syncthreads();
    // the text is read ONCE from (mapped host) memory into shared memory; both passes of the pre-tokenizer run on that copy
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(text);
        uint4 *dst = reinterpret_cast<uint4 *>(s_text);
        if constexpr (kInline) { if (tid < sizeof(InlineText) / 16) dst[tid] = inl.v[tid]; }
        else for (uint32_t v = tid; v < (n + 16 + 15) / 16; v += kSmallThreads) dst[v] = __ldg(src + v);  // (+16: the zero padding behind the text)
    }
    __syncthreads();
    if (warp < n_tiles) { const unsigned long long sum = pt::count_tile<kBert, false>(t, s_text, n, warp, s_status); if (lane == 0) s_sum[warp] = sum; }
    __syncthreads();
    if (tid == 0) {
        unsigned long long words = 0, bytes = 0;
        for (uint32_t k = 0; k < n_tiles; ++k) { const unsigned long long v = s_sum[k]; s_sum[k] = (words << 32) | bytes; words += v >> 32; bytes += v & 0xFFFFFFFFull; }
        s_status[pt::kPtWords] = (uint32_t)words;
        a.word_off[words] = (uint32_t)bytes;                                          // closing offset
    }
    __syncthreads();
    if (warp < n_tiles) pt::write_tile<kBert, false>(t, s_text, n, warp, s_sum[warp], a.arena, a.word_off, nullptr, s_status);
    __syncthreads();
    const uint32_t n_words = s_status[pt::kPtWords];
    const uint32_t clk_pretok = (uint32_t)(clock64() - clk0);
    // header + completion flag: every id store of the CTA is fenced to system scope before the flag is written
    auto finish = [&](uint32_t code, uint32_t n_tokens, uint32_t n_h6) {
        const uint32_t clk_encode = (uint32_t)(clock64() - clk0);
        __threadfence_system();
        __syncthreads();
        if (tid == 0) {
            a.out[0] = code; a.out[1] = n_tokens; a.out[2] = n_h6; a.out[3] = n_words;
            a.out[5] = clk_pretok; a.out[6] = clk_encode; a.out[7] = (uint32_t)(clock64() - clk0);
            __threadfence_system();
            *(volatile uint32_t *)(a.out + 4) = a.seq;
        }
    };
    if (s_status[pt::kPtCode] != SWT_OK) { finish(s_status[pt::kPtCode], 0u, 0u); return; }
    // ---- 2. encoder
    uint32_t h6 = 0;
    if (n_words <= (uint32_t)kSmallThreads) {
        // one word per thread, ids stay in the thread until their position is known
        uint32_t buf[kShortBytes];
        uint32_t b0 = 0, nb = 0, cnt = 0;
        if (tid < n_words) { b0 = a.word_off[tid]; nb = a.word_off[tid + 1] - b0; }
        if (!__syncthreads_or(nb > (uint32_t)kShortBytes)) {
            if (tid < n_words) cnt = enc.encode_short(nullptr, a.arena + b0, nb, buf, h6);
            if (h6) atomicAdd(&s_h6, h6);
            uint32_t total;
            const uint32_t pos = block_exclusive_scan(cnt, sh_scan, &total);
            if (pos + cnt <= a.out_cap) for (uint32_t k = 0; k < cnt; ++k) a.out[8 + pos + k] = buf[k];
            finish(total > a.out_cap ? (uint32_t)SWT_ERR_CAPACITY : (uint32_t)SWT_OK, total, s_h6);
            return;
        }
    }
    // general case.  Words of up to kShortBytes bytes: one thread each
    for (uint32_t w = tid; w < n_words; w += kSmallThreads) {
        const uint32_t b0 = a.word_off[w], nb = a.word_off[w + 1] - b0;
        if (nb > (uint32_t)kShortBytes) continue;
        uint32_t buf[kShortBytes];
        const uint32_t c = enc.encode_short(nullptr, a.arena + b0, nb, buf, h6);
        uint32_t *dst = a.scratch + b0 + w;
        for (uint32_t k = 0; k < c; ++k) dst[k] = buf[k];
        a.cnt[w] = c;
    }
    // longer words (rare): BPE by a whole warp in the ping-pong buffers, WordPiece by one lane (count, then write)
    for (uint32_t w = warp; w < n_words; w += kSmallThreads / 32) {
        const uint32_t b0 = a.word_off[w], nb = a.word_off[w + 1] - b0;
        if (nb <= (uint32_t)kShortBytes) continue;                                  // warp-uniform
        uint32_t *dst = a.scratch + b0 + w;
        if constexpr (Enc::kScratchLong) {
            uint32_t *res = nullptr, *buf_a = a.long_buf + 2 * (size_t)b0;
            const uint32_t c = enc.encode_long_warp(a.arena + b0, nb, buf_a, buf_a + nb, &res);
            for (uint32_t k = lane; k < c; k += 32) dst[k] = res[k];
            if (lane == 0) a.cnt[w] = c;
        } else if (lane == 0) {
            uint32_t dummy = 0;
            const uint32_t c = enc.long_count(a.arena + b0, nb, h6);
            enc.long_emit(a.arena + b0, nb, dst, c, dummy);
            a.cnt[w] = c;
        }
    }
    if (h6) atomicAdd(&s_h6, h6);
    __syncthreads();
    // output positions: block scan over the counts, 4 consecutive words per thread and round; then the ids move to their position
    for (uint32_t w0 = 0; w0 < n_words; w0 += 4 * kSmallThreads) {
        const uint32_t w = w0 + 4 * tid;
        uint32_t c[4], sum = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { c[k] = w + k < n_words ? a.cnt[w + k] : 0u; sum += c[k]; }
        uint32_t total;
        uint32_t pos = s_carry + block_exclusive_scan(sum, sh_scan, &total);         // (ends with a barrier: s_carry is read before it moves)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (w + k < n_words && pos + c[k] <= a.out_cap) {
                const uint32_t *src = a.scratch + a.word_off[w + k] + w + k;
                for (uint32_t i = 0; i < c[k]; ++i) a.compact[pos + i] = src[i];
            }
            pos += c[k];
        }
        if (tid == 0) s_carry += total;
        __syncthreads();
    }
    const uint32_t n_tokens = s_carry, n_copy = min(n_tokens, a.out_cap);
    uint4 *o4 = reinterpret_cast<uint4 *>(a.out + 8);
    const uint4 *c4 = reinterpret_cast<const uint4 *>(a.compact);
    for (uint32_t v = tid; v < (n_copy + 3) / 4; v += kSmallThreads) o4[v] = c4[v];
    finish(n_tokens > a.out_cap ? (uint32_t)SWT_ERR_CAPACITY : (uint32_t)SWT_OK, n_tokens, s_h6);
}
template <class Enc>
int launch_tokenize_small(const Enc &enc, const pt::PretokDev &t, bool bert, const uint8_t *h_text, const uint8_t *d_text, uint32_t n, const SmallArgs &a,
                          cudaStream_t st) {
    if (n <= kInlineTextBytes) {                       // the text rides in the parameter buffer
        InlineText inl;
        memset(&inl, 0, sizeof(inl));
        memcpy(&inl, h_text, n);
        if (bert) tokenize_small_kernel<Enc, true, true><<<1, kSmallThreads, 0, st>>>(enc, t, nullptr, inl, n, a);
        else tokenize_small_kernel<Enc, false, true><<<1, kSmallThreads, 0, st>>>(enc, t, nullptr, inl, n, a);
    } else {
        const InlineText none{};
        if (bert) tokenize_small_kernel<Enc, true, false><<<1, kSmallThreads, 0, st>>>(enc, t, d_text, none, n, a);
        else tokenize_small_kernel<Enc, false, false><<<1, kSmallThreads, 0, st>>>(enc, t, d_text, none, n, a);
    }
    SWT_CUDA_OK(cudaGetLastError());
    return SWT_OK;
}

template <class Enc>
int launch_encode_tiles(const Enc &enc, const uint8_t *d_arena, const uint32_t *d_word_off, uint32_t n_words, uint64_t long_word_bytes,
                        uint32_t *d_out_ids, uint64_t out_cap, uint32_t *d_out_tok_off, uint32_t tok_base, void *d_workspace,
                        size_t workspace_bytes, uint32_t *d_status, cudaStream_t st) {
    SWT_REQUIRE(d_word_off && d_status && d_workspace, "NULL argument");
    SWT_REQUIRE(n_words == 0 || (d_arena && d_out_ids), "NULL data pointer");
    EncodeWorkspace ws;
    size_t need = encode_workspace_layout(n_words, long_word_bytes, d_workspace, &ws);
    if (need > workspace_bytes) { set_error("encode workspace too small"); return SWT_ERR_CAPACITY; }
    SWT_CUDA_OK(cudaMemsetAsync(d_status, 0, 8 * sizeof(uint32_t), st));
    if (n_words == 0) {
        if (d_out_tok_off) SWT_CUDA_OK(cudaMemcpyAsync(d_out_tok_off, &tok_base, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        return SWT_OK;
    }
    static int grid1 = 0, grid2 = 0;
    if (!grid1) {
        grid1 = encode_grid((const void *)encode_count_kernel<Enc>, kThreads, 0);
        grid2 = encode_grid((const void *)encode_emit_kernel<Enc>, kThreads, 0);
    }
    const uint32_t n_ctas = (ws.n_tiles + kWarps - 1) / kWarps, n_groups = (ws.n_tiles + kGroupTiles - 1) / kGroupTiles;
    // swt_tune("timing", 1): per-kernel CUDA-event times of this call on stderr (diagnostic; synchronises)
    const bool timing = g_tune.timing != 0;
    cudaEvent_t ev[6];
    if (timing) for (auto &e : ev) cudaEventCreate(&e);
    if (timing) cudaEventRecord(ev[0], st);
    SWT_CUDA_OK(cudaMemsetAsync(ws.long_tiles, 0, ((size_t)(ws.n_tiles + 31) / 32) * 4, st));
    memo_clear_kernel<<<kNumSMs * 8, 256, 0, st>>>(ws.keys, ws.memo_mask ? ws.memo_mask + 1 : 0u, ws.long_cursor);
    if (timing) cudaEventRecord(ev[1], st);
    const uint32_t warp_words = (uint32_t)std::max(g_tune.warp_words, 0);
    if (ws.memo_mask == 0) {                           // memo disabled: the direct path, one lane per word, no queue
        static int grid_direct = 0;
        if (!grid_direct) grid_direct = encode_grid((const void *)encode_count_direct_kernel<Enc>, kThreads, 0);
        encode_count_direct_kernel<Enc><<<(int)std::min<uint32_t>((uint32_t)grid_direct, n_ctas), kThreads, 0, st>>>(enc, d_arena, d_word_off, n_words, ws, d_status);
    } else {
    // warm-up: the first tiles go through the kernel that resolves its misses itself (the memo fills with the frequent types) ...
    const uint32_t kWarmTiles = (uint32_t)std::max(1, g_tune.split_warm_tiles);     // default 1 M words
    const bool split = Enc::kSplitCount && g_tune.split_count && ws.memo_mask != 0 && ws.n_tiles > 4 * kWarmTiles;
    const uint32_t warm_end = split ? kWarmTiles : ws.n_tiles;
    encode_count_kernel<Enc><<<(int)std::min<uint32_t>((uint32_t)grid1, (warm_end + kWarps - 1) / kWarps), kThreads, 0, st>>>(
        enc, d_arena, d_word_off, n_words, ws, d_status, warp_words, 0u, warm_end, split ? 1u : 0u, kWarmTiles);
    if (split) {
        // ... the rest in chunks: leaf count kernel (probe only), then the resolve kernel for the words it left behind.  A stream whose
        // warm-up sent more than 1/8 of its words to the slow path has MANY word types: deferring its first occurrences to the end of
        // a chunk would send every repetition inside the chunk to the slow path too (measured: 2 M types, 5.4 M -> 11.9 M slow words,
        // 91 -> 76 GB/s),
        // so the leaf / resolve kernels return at once (device-side gate, no host synchronisation) and the kernel that resolves its
        // misses itself takes the rest of the stream.
        static int grid_leaf = 0, grid_res = 0;
        if (!grid_leaf) {
            grid_leaf = encode_grid((const void *)encode_count_leaf_kernel, kThreads, 0);
            grid_res = encode_grid((const void *)encode_resolve_kernel<Enc>, kThreads, 0);
        }
        SWT_CUDA_OK(cudaMemsetAsync(ws.tile_pend, 0, (size_t)ws.n_tiles * sizeof(uint2), st));
        const uint32_t kChunks = (uint32_t)std::max(1, g_tune.split_chunks);
        const uint32_t per = ((ws.n_tiles - warm_end + kChunks - 1) / kChunks + 31u) & ~31u;
        for (uint32_t t0 = warm_end; t0 < ws.n_tiles; t0 += per) {
            const uint32_t t1 = std::min(ws.n_tiles, t0 + per), nt = t1 - t0;
            encode_count_leaf_kernel<<<(int)std::min<uint32_t>((uint32_t)grid_leaf, (nt + kWarps - 1) / kWarps), kThreads, 0, st>>>(
                d_arena, d_word_off, n_words, ws, d_status, t0, t1, kWarmTiles);
            encode_resolve_kernel<Enc><<<(int)std::min<uint32_t>((uint32_t)grid_res, (nt / 32 + kWarps) / kWarps), kThreads, 0, st>>>(
                enc, d_arena, d_word_off, n_words, ws, d_status, warp_words, t0, t1, kWarmTiles);
        }
        encode_count_kernel<Enc><<<(int)std::min<uint32_t>((uint32_t)grid1, (ws.n_tiles - warm_end + kWarps - 1) / kWarps), kThreads, 0, st>>>(
            enc, d_arena, d_word_off, n_words, ws, d_status, warp_words, warm_end, ws.n_tiles, 2u, kWarmTiles);
    }
    }
    const int long_grid = (int)std::min<uint32_t>(kNumSMs * 2, ((ws.n_tiles + 31) / 32 + 7) / 8);      // one warp per 32 tiles of the bitmap
    encode_long_count_kernel<Enc><<<long_grid, 256, 0, st>>>(enc, d_arena, d_word_off, n_words, ws, d_status);
    if (timing) cudaEventRecord(ev[2], st);
    encode_scan_groups_kernel<<<n_groups, 256, 0, st>>>(ws);
    if (timing) cudaEventRecord(ev[3], st);
    encode_scan_top_kernel<<<1, 1024, 0, st>>>(ws, n_words, d_out_tok_off, tok_base, out_cap, d_status);
    if (timing) cudaEventRecord(ev[4], st);
    encode_emit_kernel<Enc><<<(int)std::min<uint32_t>((uint32_t)grid2, n_ctas), kThreads, 0, st>>>(enc, d_arena, d_word_off, n_words, d_out_ids, out_cap,
                                                                                                   d_out_tok_off, tok_base, ws);
    encode_long_emit_kernel<Enc><<<long_grid, 256, 0, st>>>(enc, d_arena, d_word_off, n_words, d_out_ids, out_cap, ws);
    if (timing) {
        cudaEventRecord(ev[5], st); cudaEventSynchronize(ev[5]);
        float t[5];
        for (int i = 0; i < 5; ++i) cudaEventElapsedTime(&t[i], ev[i], ev[i + 1]);
        fprintf(stderr, "[swt timing] clear %.3f count %.3f scan %.3f+%.3f emit %.3f ms (grid %d/%d, %u tiles, memo %u slots)\n", t[0], t[1], t[2], t[3],
                t[4], grid1, grid2, ws.n_tiles, ws.memo_mask ? ws.memo_mask + 1 : 0u);
        for (auto &e : ev) cudaEventDestroy(e);
    }
    SWT_CUDA_OK(cudaGetLastError());
    return SWT_OK;
}
#endif  // __CUDACC__

}  // namespace swt
