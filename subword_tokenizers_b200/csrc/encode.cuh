// encode.cuh -- the tile kernels shared by all encoders (HP-1 FastBPE, HP-2 FastWP, and the NaiveBPE / NaiveWP encoders).
//
// Data flow of one encode call (north-star subsystem 1: packed word-offset/byte arena).  Tile = 64 consecutive words per
// WARP (2 per lane), tiles assigned round-robin to a persistent grid; warps never synchronise with each other.
//
//   pass 1  encode_count_kernel   per word: look the word up in the word-type memo; on a miss encode it (rank table /
//                                 trie walk) and publish the ids.  Writes one packed u32 record per word (kind, token
//                                 count, memo slot) and the tile's token total.
//   scan    two tiny kernels      exclusive scan of the tile totals (in-group prefixes + group bases)
//   pass 2  encode_emit_kernel    per word: copy the ids memo entry -> the warp's shared-memory buffer at their tile-local
//                                 position; the tile's ids leave shared memory as 128-bit coalesced stores at their final
//                                 position, u32 token offsets as 8-byte stores.
//
// An earlier single-pass version (decoupled look-back over tile states) was measured slower: with warp-sized tiles
// the look-back walked hundreds of in-flight predecessors, with CTA-sized tiles the CTA barriers serialised the L2
// latencies.  The two passes cost 8 extra bytes of HBM traffic per word and have no inter-tile dependency at all.
//
// Word-type memo: encode_word is a pure function of the word and word streams are
// Zipf-distributed, so every launch keeps a hash table  word bytes -> token ids  in its workspace.  The first
// thread that meets a word type claims a slot with ONE 128-bit CAS (key = first 15 bytes + length; the
// remaining bytes of longer words are stored in the entry and compared, so a hit is always exact), encodes
// the word and publishes the ids with a release store; later occurrences copy the ids instead of re-walking
// the rank table / trie.  The memo is rebuilt from empty by every launch (nothing is carried over between
// calls), readers never wait (a slot that is claimed but not yet published is simply recomputed), and words
// that do not fit (longer than 32 bytes, table full) take the direct path.
// Every input byte is still read and every output id still written by every launch.
#pragma once
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace swt {

constexpr int kWarps = 8;              // warps per CTA; every warp owns its own tile
constexpr int kThreads = kWarps * 32;
constexpr int kCtasPerSm = 4;
constexpr int kEmitCtasPerSm = 4;
constexpr int kWordsPerThread = 2;
constexpr int kTileWords = 32 * kWordsPerThread;         // 64 words per (warp) tile
constexpr int kShortBytes = 32;        // words up to this many bytes are encoded by one thread (and memoised)
constexpr int kCompactTokens = 640;    // tile token totals up to this are assembled in smem before the store
constexpr int kMemoTokens = 44;        // >= kShortBytes: every word of up to 32 bytes fits (ids <= bytes)
constexpr int kMemoProbes = 8;

// status words written by the encode kernels
enum { kStatusCode = 0, kStatusTokens = 1, kStatusH6 = 2, kStatusTokensHi = 3, kStatusMemoTypes = 4, kStatusSlowWords = 5 };

struct alignas(256) MemoEntry {         // 256 bytes; a fast-path hit touches bytes 0..15 (count pass) and 32..47 (emit pass)
    unsigned long long lo, hi;          // CAS key: word bytes 0..7 | bytes 8..14, bits 56-59 = length (0 for words > 15 bytes),
                                        // bits 60-63 = "pub": 0, or n_tokens + 1 once ids16[] is valid (set after the CAS claim)
    uint32_t meta;                      // 0 = claimed, not published; else (n_tokens + 1) | (h6 << 8); ~0 = not cacheable
    uint32_t tail_last;                 // words > 15 bytes: byte 31 | length << 8
    uint32_t pad[2];                    // (bytes 0..31 = the one sector memo_clear_kernel has to zero)
    uint16_t ids16[16];                 // at byte 32: ids 0..15 as 16-bit values (only when every id of the word is < 65536)
    unsigned long long tail_a, tail_b;  // at byte 64; words > 15 bytes: bytes 15..22 | 23..30 (verified after the key)
    uint32_t tok[kMemoTokens];          // all ids as 32-bit values, at byte 80
};
constexpr uint32_t kMemoSlotBits = 23, kMemoSlotMask = (1u << kMemoSlotBits) - 1u;     // slot field of the per-word record
static_assert(sizeof(MemoEntry) == 256, "MemoEntry must be 256 bytes");
constexpr unsigned long long kPubMask = 0xFull << 60;

struct MemoKey { unsigned long long lo, hi, tail_a, tail_b; uint32_t tail_last; uint32_t nbytes; };

struct EncodeWorkspace {
    unsigned long long *long_cursor; // 1 (BPE: allocation cursor into long_scratch, in u32 units)
    MemoEntry *memo; uint32_t memo_mask;   // memo_mask == 0: memo disabled
    uint32_t *packed;                // n_words: per-word record handed from the count pass to the emit pass
    uint32_t *tile_total;            // n_tiles: tokens per tile, then (after the scan) the in-group exclusive prefix
    unsigned long long *group_base;  // n_groups: tokens per group of 1024 tiles, then the group's global token offset
    uint32_t *long_scratch;          // BPE symbol ping-pong buffers for long words
    uint64_t long_scratch_elems;
    uint32_t n_tiles;
    size_t zero_bytes;               // prefix of the workspace that must be zeroed before a launch
};

size_t encode_workspace_layout(uint32_t n_words, uint64_t long_bytes, void *base, EncodeWorkspace *ws);
int encode_grid(const void *kernel, int block, size_t dyn_smem);

#ifdef __CUDACC__
enum { kMemoHit = 0, kMemoClaimed = 1, kMemoMiss = 2, kMemoPending = 3 };

__device__ __forceinline__ void cas128(MemoEntry *e, unsigned long long lo, unsigned long long hi,
                                       unsigned long long &old_lo, unsigned long long &old_hi) {
    asm volatile("{\n\t.reg .b128 c, v, o;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 v, {%4, %5};\n\t"
                 "atom.global.cas.b128 o, [%6], c, v;\n\tmov.b128 {%0, %1}, o;\n\t}"
                 : "=l"(old_lo), "=l"(old_hi) : "l"(0ull), "l"(0ull), "l"(lo), "l"(hi), "l"(e) : "memory");
}
// the memo is written during the launch: read it at L2 (never through the non-coherent L1)
__device__ __forceinline__ void ld_cg_u64x2(const void *p, unsigned long long &a, unsigned long long &b) {
    asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ uint2 ld_cg_u32x2(const void *p) {
    uint2 v;
    asm volatile("ld.global.cg.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
// L1-cached variants for the FAST path.  Hot word types (Zipf) then hit the 100+ KB L1 instead of going to L2.  This is
// safe although the memo is written during the launch: a key never changes once set and meta goes 0 -> final exactly
// once, after the ids of the same sector were written (release); L1 fills are whole 32-byte sectors.  A stale L1 sector can
// therefore only show "empty" or "not published yet", which sends the word to the slow path, and the slow path re-probes
// at L2 (ld.cg).  Id sectors beyond the first are only ever read after a valid meta was seen, i.e. after publication.
__device__ __forceinline__ void ld_ca_u64x2(const void *p, unsigned long long &a, unsigned long long &b) {
    asm volatile("ld.global.ca.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ uint4 ld_ca_u32x4(const void *p) {
    uint4 v;
    asm volatile("ld.global.ca.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_ca_u32(const void *p) {
    uint32_t v;
    asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// (no acquire loads anywhere: ld.acquire.gpu compiles to LDG + CCTL.IVALL, a full L1 invalidate per probe.  Ordering comes from the
// writer's release -- ids are performed at L2 before meta / the pub nibble -- plus the reader's control dependency.)
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

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

// slot hash of a key, in 32-bit operations (the 64-bit mixer costs ~4x the instructions on the fast path)
__device__ __forceinline__ uint32_t memo_hash4(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
    uint32_t h = (w0 ^ 0x9E3779B9u) * 0x85EBCA6Bu;
    h = (h ^ (h >> 15) ^ w1) * 0xC2B2AE35u;
    h = (h ^ (h >> 13) ^ w2) * 0x27D4EB2Fu;
    h = (h ^ (h >> 16) ^ w3) * 0x165667B1u;
    return h ^ (h >> 15);
}
__device__ __forceinline__ uint32_t memo_hash(unsigned long long lo, unsigned long long hi) {
    return memo_hash4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
}

// Probes the memo at L2 (slow path).  kMemoHit: slot/meta describe a published entry for exactly this word.
// kMemoClaimed: this thread now owns `slot` and must call memo_publish after encoding.  kMemoPending: another thread owns the
// slot of exactly this word (words of up to 15 bytes: the key is the whole word) and will have published it by the time the
// emit pass runs -- encode for the count, but let the emit pass read the ids from `slot`.  kMemoMiss: encode directly, publish nothing.
static __device__ __noinline__ int memo_probe(const EncodeWorkspace &ws, const MemoKey &key, uint32_t &slot, uint32_t &meta_out) {
    uint32_t h = memo_hash(key.lo, key.hi) & ws.memo_mask;
    for (int probe = 0; probe < kMemoProbes; ++probe, h = (h + 1) & ws.memo_mask) {
        MemoEntry *e = ws.memo + h;
        unsigned long long klo, khi;
        ld_cg_u64x2(e, klo, khi);
        uint2 m = ld_cg_u32x2(&e->meta);                                 // meta, tail_last
        if (klo == 0 && khi == 0) {
            cas128(e, key.lo, key.hi, klo, khi);
            if (klo == 0 && khi == 0) { slot = h; return kMemoClaimed; }
            m.x = 0;                                                     // lost the race: the winner has not published yet
        }
        if (klo != key.lo || (khi & ~kPubMask) != key.hi) continue;
        if (m.x == 0 && key.nbytes <= 15) { slot = h; return kMemoPending; }   // claimed by another thread, not published yet
        if (m.x == 0 || m.x == 0xFFFFFFFFu) return kMemoMiss;            // (long word: tail not comparable yet) / not cacheable
        if (key.nbytes > 15) {                                           // same 15-byte prefix: check the rest and the length
            unsigned long long ta, tb;
            ld_cg_u64x2(&e->tail_a, ta, tb);
            if (ta != key.tail_a || tb != key.tail_b || m.y != key.tail_last) continue;
        }
        slot = h; meta_out = m.x;
        return kMemoHit;
    }
    return kMemoMiss;
}
// returns true when the entry now holds the ids (false: too many ids, marked not cacheable)
__device__ __forceinline__ bool memo_publish(const EncodeWorkspace &ws, uint32_t slot, const MemoKey &key, const uint32_t *buf,
                                             uint32_t ntok, uint32_t h6, uint32_t *status) {
    MemoEntry *e = ws.memo + slot;
    if (ntok > (uint32_t)kMemoTokens || h6 > 0xFFFFFFu) { st_release_u32(&e->meta, 0xFFFFFFFFu); return false; }
    e->tail_a = key.tail_a; e->tail_b = key.tail_b; e->tail_last = key.tail_last;
    bool narrow = ntok <= 14 && h6 == 0;                                // servable by the one-load fast path?
    for (uint32_t k = 0; k < ntok; ++k) { e->tok[k] = buf[k]; narrow = narrow && buf[k] < 65536u; }
    if (narrow) for (uint32_t k = 0; k < ntok; ++k) e->ids16[k] = (uint16_t)buf[k];
    st_release_u32(&e->meta, (ntok + 1) | (h6 << 8));                   // release: everything above is visible first
    if (narrow) atomicOr(reinterpret_cast<uint32_t *>(&e->hi) + 1, (ntok + 1) << 28);   // pub nibble, after the release
    atomicAdd(&status[kStatusMemoTypes], 1u);
    return true;
}

// per-word record between the two passes, packed into 32 bits:
//   [31:29] kind; Hit / Hit16 (ids16[] valid): [28:23] n_tokens, [22:0] memo slot; Recompute: [28:23] n_tokens;
//   WP long: [28:0] n_tokens; BPE long: [28:0] scratch granule (16 u32) -- header word 0 holds n_tokens
enum : uint32_t { kWordNone = 0u, kWordHit = 1u, kWordRecompute = 2u, kWordLong = 3u, kWordLongB = 4u, kWordHit16 = 5u };
constexpr uint32_t kLongHeader = 16;      // u32 words reserved in front of the two scratch buffers of a long BPE word
constexpr uint32_t kGroupTiles = 1024;    // tiles per scan group

struct SlowResult { uint32_t kind, ntok, slot, h6; };

// pass 1 slow path (first probe did not hit): full probe at L2, then direct encode + publish.  Out of line so that
// the unrolled fast path stays small.
template <class Enc>
__device__ __noinline__ SlowResult resolve_slow(const Enc &enc, const EncodeWorkspace &ws, const uint8_t *arena, uint32_t b0,
                                                uint32_t nbytes, uint32_t arena_end, uint32_t *status) {
    uint32_t buf[kShortBytes];                                   // scratch for one directly encoded word
    SlowResult r; r.kind = kWordRecompute; r.ntok = 0; r.slot = 0; r.h6 = 0;
    int m = kMemoMiss; uint32_t meta = 0; MemoKey key;
    if (ws.memo_mask && nbytes >= 1) { memo_key(arena, b0, nbytes, arena_end, key); m = memo_probe(ws, key, r.slot, meta); }
    if (m == kMemoHit) { r.kind = kWordHit; r.ntok = (meta & 0xFFu) - 1; r.h6 = meta >> 8; return r; }
    r.ntok = enc.encode_short(arena + b0, nbytes, buf, r.h6);
    if (m == kMemoClaimed && memo_publish(ws, r.slot, key, buf, r.ntok, r.h6, status)) r.kind = kWordHit;
    // the owner publishes under the same condition (memo_publish), so the ids will be there for the emit pass
    if (m == kMemoPending && r.ntok <= (uint32_t)kMemoTokens && r.h6 <= 0xFFFFFFu) r.kind = kWordHit;
    return r;
}
// pass 2 slow paths: re-encode a word whose ids were not kept, or emit a long word (Enc without scratch)
template <class Enc>
__device__ __noinline__ void emit_slow(const Enc &enc, const uint8_t *word, uint32_t nbytes, uint32_t kind, uint32_t ntok,
                                       uint32_t *dst) {
    if (kind == kWordRecompute) {
        uint32_t buf[kShortBytes];
        uint32_t dummy = 0;
        const uint32_t n = enc.encode_short(word, nbytes, buf, dummy);
        for (uint32_t k = 0; k < n; ++k) dst[k] = buf[k];
    } else if constexpr (!Enc::kScratchLong) {
        uint32_t dummy = 0;
        enc.long_emit(word, nbytes, dst, ntok, dummy);
    }
}

// ids of a one-load hit (16-bit ids, at most 14) -> dst.  The first eight are predicated stores without branches
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
template <bool kShared>
__device__ __forceinline__ void store_hit16_ids(uint32_t *dst, uint32_t n, uint4 a, uint4 b) {
    store_id_if<kShared>(dst, 0, n, a.x & 0xFFFFu); store_id_if<kShared>(dst, 1, n, a.x >> 16);
    store_id_if<kShared>(dst, 2, n, a.y & 0xFFFFu); store_id_if<kShared>(dst, 3, n, a.y >> 16);
    store_id_if<kShared>(dst, 4, n, a.z & 0xFFFFu); store_id_if<kShared>(dst, 5, n, a.z >> 16);
    store_id_if<kShared>(dst, 6, n, a.w & 0xFFFFu); store_id_if<kShared>(dst, 7, n, a.w >> 16);
    if (n > 8) {
        store_id_if<kShared>(dst, 8, n, b.x & 0xFFFFu); store_id_if<kShared>(dst, 9, n, b.x >> 16);
        store_id_if<kShared>(dst, 10, n, b.y & 0xFFFFu); store_id_if<kShared>(dst, 11, n, b.y >> 16);
        store_id_if<kShared>(dst, 12, n, b.z & 0xFFFFu); store_id_if<kShared>(dst, 13, n, b.z >> 16);
    }
}
// ids of a hit found by the slow path (32-bit id list; ids 0-7 were prefetched, the rest is fetched here)
__device__ __forceinline__ void store_hit_ids(uint32_t *dst, uint32_t n, uint4 a, uint4 b, const MemoEntry *e) {
    if (n > 0) dst[0] = a.x;
    if (n > 1) dst[1] = a.y;
    if (n > 2) dst[2] = a.z;
    if (n > 3) dst[3] = a.w;
    if (n > 4) dst[4] = b.x;
    if (n > 5) dst[5] = b.y;
    if (n > 6) dst[6] = b.z;
    if (n > 7) dst[7] = b.w;
    for (uint32_t k0 = 8; k0 < n; k0 += 4) {
        const uint4 v = ld_ca_u32x4(&e->tok[k0]);
        dst[k0] = v.x;
        if (k0 + 1 < n) dst[k0 + 1] = v.y;
        if (k0 + 2 < n) dst[k0 + 2] = v.z;
        if (k0 + 3 < n) dst[k0 + 3] = v.w;
    }
}

// clears key and meta of every memo entry (the rest of an entry is only read after its meta / pub nibble was published)
static __global__ void __launch_bounds__(256) memo_clear_kernel(MemoEntry *memo, uint32_t n_slots, unsigned long long *long_cursor) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += gridDim.x * blockDim.x) {
        uint4 *e = reinterpret_cast<uint4 *>(memo + i);
        e[0] = z; e[1] = z;                                 // key (bytes 0-15) and meta / tail_last (bytes 16-31): one sector
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *long_cursor = 0ull;
}

// resolves `count` (<= 32) pending words of a warp, one per lane: record, token count into the tile total.  Returns the lane's
// H6 events.  Out of line: keeps the registers of the slow path out of the count loop.
template <class Enc>
__device__ __noinline__ uint32_t flush_pending_words(const Enc &enc, const EncodeWorkspace &ws, const uint8_t *arena, const uint32_t *word_off,
                                                     uint32_t arena_end, uint32_t *status, const uint32_t *pend, uint32_t count) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t h6 = 0;
    if (lane < count) {
        const uint32_t w = pend[lane];
        const uint32_t wb0 = __ldg(word_off + w), wnb = __ldg(word_off + w + 1) - wb0;
        const SlowResult r = resolve_slow(enc, ws, arena, wb0, wnb, arena_end, status);
        ws.packed[w] = (r.kind << 29) | (r.ntok << 23) | (r.kind == kWordHit ? r.slot : 0u);
        if (r.ntok) atomicAdd(&ws.tile_total[w / kTileWords], r.ntok);
        h6 = r.h6;
    }
    __syncwarp();
    return h6;
}

// ---- pass 1: count ---------------------------------------------------------------------------------------------------
// Enc provides
//   uint32_t encode_short(const uint8_t *p, uint32_t nbytes, uint32_t *buf /*thread-local, kShortBytes*/, uint32_t &h6) const
//   static constexpr bool kScratchLong, kBatchSlowPath
//   kScratchLong == false: uint32_t long_count(p, nbytes, h6) const;  void long_emit(p, nbytes, uint32_t *dst, uint32_t cap, uint32_t &h6) const
//   kScratchLong == true : uint32_t encode_long_warp(p, nbytes, bufA, bufB, uint32_t **result) const   (all 32 lanes)
//
// Every warp owns tiles of kTileWords = 64 consecutive words (2 per lane), assigned round-robin; warps never wait for
// each other, so the L2 latencies of one warp's probes are covered by the other warps of the SM.
template <class Enc>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
encode_count_kernel(Enc enc, const uint8_t *__restrict__ arena, const uint32_t *__restrict__ word_off, uint32_t n_words,
                    EncodeWorkspace ws, uint32_t *status) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * kThreads + threadIdx.x) >> 5, n_warps = (gridDim.x * kThreads) >> 5;
    const uint32_t arena_end = word_off[n_words];
    const bool use_memo = ws.memo_mask != 0;
    uint32_t h6 = 0;

    // offsets of this lane's two words in a tile (three consecutive offsets); the next tile's are fetched one tile ahead
    auto load_offsets = [&](uint32_t t, uint32_t &o0, uint32_t &o1, uint32_t &o2) {
        const uint32_t w0 = t * kTileWords, tw = min((uint32_t)kTileWords, n_words - w0), i0 = lane * kWordsPerThread;
        o0 = i0 <= tw ? __ldg(word_off + w0 + i0) : 0u;
        o1 = i0 + 1 <= tw ? __ldg(word_off + w0 + i0 + 1) : 0u;
        o2 = i0 + 2 <= tw ? __ldg(word_off + w0 + i0 + 2) : 0u;
    };
    // pending queue of the warp: words waiting for the slow path (at most 31 left over + 64 from one tile)
    __shared__ uint32_t s_pend[kWarps][96];
    uint32_t *pend = s_pend[threadIdx.x >> 5];
    uint32_t n_pend = 0;
    uint32_t n_slow_words = 0;                                                      // diagnostic (status word 5), warp-uniform
    auto flush_pending = [&](uint32_t first, uint32_t count) {
        h6 += flush_pending_words(enc, ws, arena, word_off, arena_end, status, pend + first, count);
        n_slow_words += count;
    };
    uint32_t po0 = 0, po1 = 0, po2 = 0;
    if (warp_global < ws.n_tiles) load_offsets(warp_global, po0, po1, po2);
    for (uint32_t tile = warp_global; tile < ws.n_tiles; tile += n_warps) {
        const uint32_t w_tile = tile * kTileWords;
        const uint32_t tile_words = min((uint32_t)kTileWords, n_words - w_tile);
        uint32_t nb[kWordsPerThread], b0s[kWordsPerThread];
        {
            const uint32_t i0 = lane * kWordsPerThread;
            const uint32_t o0 = po0, o1 = po1, o2 = po2;
            if (tile + n_warps < ws.n_tiles) load_offsets(tile + n_warps, po0, po1, po2);
            b0s[0] = o0; nb[0] = i0 < tile_words ? o1 - o0 : 0xFFFFFFFFu;       // 0xFFFFFFFF: no word
            b0s[1] = o1; nb[1] = i0 + 1 < tile_words ? o2 - o1 : 0xFFFFFFFFu;
        }
        // ---- fast path (words of 1..15 bytes): 128-bit key from aligned 8-byte loads, then up to two L1-cached memo
        // probes.  Both words' loads are in flight together.
        uint32_t kind[kWordsPerThread], ntok[kWordsPerThread], slot[kWordsPerThread];
        uint4 kw[kWordsPerThread], ew[kWordsPerThread];           // key of the word / key found in the probed entry
        bool fastj[kWordsPerThread], slow[kWordsPerThread], is_long[kWordsPerThread];
        {
            uint32_t a0[kWordsPerThread], a1[kWordsPerThread], a2[kWordsPerThread], a3[kWordsPerThread], a4[kWordsPerThread];
#pragma unroll
            for (int j = 0; j < kWordsPerThread; ++j) {
                kind[j] = kWordNone; ntok[j] = 0; slot[j] = 0; slow[j] = false;
                kw[j] = ew[j] = make_uint4(0, 0, 0, 0);
                is_long[j] = nb[j] != 0xFFFFFFFFu && nb[j] > (uint32_t)kShortBytes;
                // words of 16..32 bytes take the same first probe on their 15-byte prefix (length nibble 0); their tail is checked below
                fastj[j] = use_memo && nb[j] >= 1 && nb[j] <= (uint32_t)kShortBytes && (uint64_t)b0s[j] + 40 <= arena_end;
                a0[j] = a1[j] = a2[j] = a3[j] = a4[j] = 0;
                if (fastj[j]) {                                   // the (up to five) aligned 32-bit words that hold the word
                    const uintptr_t a = (uintptr_t)(arena + b0s[j]);
                    const uint32_t *q = (const uint32_t *)(a & ~(uintptr_t)3);
                    const uint32_t span = min(nb[j], 15u) + (uint32_t)(a & 3);
                    a0[j] = __ldg(q);
                    a1[j] = ldg_u32_if(q + 1, span > 4); a2[j] = ldg_u32_if(q + 2, span > 8);
                    a3[j] = ldg_u32_if(q + 3, span > 12); a4[j] = ldg_u32_if(q + 4, span > 16);
                }
            }
#pragma unroll
            for (int j = 0; j < kWordsPerThread; ++j) {
                if (fastj[j]) {
                    const uint32_t sh = (uint32_t)((uintptr_t)(arena + b0s[j]) & 3) * 8, n = min(nb[j], 15u);
                    uint32_t w0 = __funnelshift_r(a0[j], a1[j], sh), w1 = __funnelshift_r(a1[j], a2[j], sh);
                    uint32_t w2 = __funnelshift_r(a2[j], a3[j], sh), w3 = __funnelshift_r(a3[j], a4[j], sh);
                    // zero the bytes at and beyond n, put the length into the top byte (layout of MemoEntry::lo/hi)
                    w0 &= low_bytes_mask((int)n); w1 &= low_bytes_mask((int)n - 4); w2 &= low_bytes_mask((int)n - 8);
                    w3 = (w3 & low_bytes_mask((int)n - 12)) | ((nb[j] <= 15u ? n : 0u) << 24);
                    kw[j] = make_uint4(w0, w1, w2, w3);
                    slot[j] = memo_hash4(w0, w1, w2, w3) & ws.memo_mask;
                    ew[j] = ld_ca_u32x4(ws.memo + slot[j]);      // ONE scattered load per word: key + pub nibble
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kWordsPerThread; ++j) {
            if (nb[j] == 0xFFFFFFFFu) continue;
            if (is_long[j]) { kind[j] = kWordLong; continue; }
            bool hit = false;
            if (fastj[j]) {
                bool same = ew[j].x == kw[j].x && ew[j].y == kw[j].y && ew[j].z == kw[j].z && ((ew[j].w ^ kw[j].w) & 0x0FFFFFFFu) == 0;
                if (!same && (ew[j].x | ew[j].y | ew[j].z | ew[j].w) != 0) {
                    // the slot holds another word (hash collision): look at the next slot
                    slot[j] = (slot[j] + 1) & ws.memo_mask;
                    ew[j] = ld_ca_u32x4(ws.memo + slot[j]);
                    same = ew[j].x == kw[j].x && ew[j].y == kw[j].y && ew[j].z == kw[j].z && ((ew[j].w ^ kw[j].w) & 0x0FFFFFFFu) == 0;
                }
                hit = same && (ew[j].w >> 28) != 0;                  // pub nibble: ids16[] valid, n_tokens + 1
                if (hit && nb[j] > 15u) {                            // rare: compare bytes 15.. with the tail stored in the entry
                    MemoKey key;
                    memo_key(arena, b0s[j], nb[j], arena_end, key);
                    const MemoEntry *e = ws.memo + slot[j];
                    unsigned long long ta, tb;
                    ld_ca_u64x2(&e->tail_a, ta, tb);
                    hit = ta == key.tail_a && tb == key.tail_b && ld_ca_u32(&e->tail_last) == key.tail_last;
                }
            }
            if (hit) { kind[j] = kWordHit16; ntok[j] = (ew[j].w >> 28) - 1; }
            else slow[j] = true;
        }
        // Words not served by the first probe (first occurrence, hash collision, more than 14 tokens).  A single trie walk /
        // merge loop is a chain of dependent L2 accesses (10-25 us) during which the other lanes of the warp would idle.
        // kBatchSlowPath: they go to the warp's pending queue and are resolved 32 at a time, one per lane (flush below) --
        // 5 % faster on the bench stream for FastWP and 30-40 % on streams with 10^5..10^6 word types.  Otherwise (FastBPE,
        // where the queue cost the count pass 10 %): the slow words of the tile are spread over the lanes and resolved now.
        if constexpr (Enc::kBatchSlowPath) {
#pragma unroll
            for (int j = 0; j < kWordsPerThread; ++j) {
                const uint32_t m = __ballot_sync(0xffffffffu, slow[j]);
                if (m == 0) continue;                                               // warp-uniform
                if (slow[j]) pend[n_pend + __popc(m & ((1u << lane) - 1u))] = w_tile + lane * kWordsPerThread + j;
                n_pend += __popc(m);
            }
        } else {
#pragma unroll
            for (int j = 0; j < kWordsPerThread; ++j) {
                const uint32_t m = __ballot_sync(0xffffffffu, slow[j]);
                if (m == 0) continue;                                               // warp-uniform
                const uint32_t n_slow = __popc(m);
                const uint32_t src = __fns(m, 0, lane + 1);                         // lane -> owner of the lane-th slow word
                const uint32_t sb0 = __shfl_sync(0xffffffffu, b0s[j], src & 31), snb = __shfl_sync(0xffffffffu, nb[j], src & 31);
                SlowResult r; r.kind = kWordNone; r.ntok = 0; r.slot = 0; r.h6 = 0;
                if (lane < n_slow) r = resolve_slow(enc, ws, arena, sb0, snb, arena_end, status);
                h6 += r.h6;
                const uint32_t rank = __popc(m & ((1u << lane) - 1u));             // this lane's word was resolved by lane `rank`
                const uint32_t k_ = __shfl_sync(0xffffffffu, r.kind, rank), n_ = __shfl_sync(0xffffffffu, r.ntok, rank);
                const uint32_t s_ = __shfl_sync(0xffffffffu, r.slot, rank);
                if (slow[j]) { kind[j] = k_; ntok[j] = n_; slot[j] = s_; slow[j] = false; }
                n_slow_words += n_slow;
            }
        }
        // long words (rare)
        uint32_t packed[kWordsPerThread];
#pragma unroll
        for (int j = 0; j < kWordsPerThread; ++j) {
            packed[j] = (kind[j] << 29) | (ntok[j] << 23) | ((kind[j] == kWordHit || kind[j] == kWordHit16) ? slot[j] : 0u);
            if constexpr (!Enc::kScratchLong) {
                if (is_long[j]) { ntok[j] = enc.long_count(arena + b0s[j], nb[j], h6); packed[j] = (kWordLong << 29) | ntok[j]; }
            } else {
                uint32_t m = __ballot_sync(0xffffffffu, is_long[j]);
                while (m) {                                                         // the whole warp works on one long word
                    const uint32_t owner = __ffs(m) - 1; m &= m - 1;
                    const uint32_t lb0 = __shfl_sync(0xffffffffu, b0s[j], owner), lnb = __shfl_sync(0xffffffffu, nb[j], owner);
                    const unsigned long long need = (kLongHeader + 2ull * lnb + 15ull) & ~15ull;
                    unsigned long long so = 0;
                    if (lane == 0) so = atomicAdd(ws.long_cursor, need);
                    so = __shfl_sync(0xffffffffu, so, 0);
                    uint32_t c = 0; uint32_t *res = nullptr;
                    const bool fits = so + need <= ws.long_scratch_elems && (so >> 4) < (1ull << 29);
                    uint32_t *bufA = ws.long_scratch + so + kLongHeader;
                    if (fits) c = enc.encode_long_warp(arena + lb0, lnb, bufA, bufA + lnb, &res);
                    else if (lane == 0) atomicExch(&status[kStatusCode], (uint32_t)SWT_ERR_CAPACITY);
                    if (lane == owner) {
                        ntok[j] = c;
                        if (fits) { ws.long_scratch[so] = c; packed[j] = ((res == bufA ? kWordLong : kWordLongB) << 29) | (uint32_t)(so >> 4); }
                        else packed[j] = 0;
                    }
                }
            }
        }
        // ---- per-word records and the tile total (pending words: record and token count are added by the flush)
        {
            const uint32_t i0 = lane * kWordsPerThread;
            if (i0 + 1 < tile_words && !slow[0] && !slow[1]) *reinterpret_cast<uint2 *>(ws.packed + w_tile + i0) = make_uint2(packed[0], packed[1]);
            else {
                if (i0 < tile_words && !slow[0]) ws.packed[w_tile + i0] = packed[0];
                if (i0 + 1 < tile_words && !slow[1]) ws.packed[w_tile + i0 + 1] = packed[1];
            }
        }
        uint32_t total = ntok[0] + ntok[1];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
        if (lane == 0) ws.tile_total[tile] = total;
        if constexpr (Enc::kBatchSlowPath) {
            __syncwarp();
            while (n_pend >= 32) { n_pend -= 32; flush_pending(n_pend, 32); }
        }
    }
    if constexpr (Enc::kBatchSlowPath) { if (n_pend) flush_pending(0, n_pend); }
    if (h6) atomicAdd(&status[kStatusH6], h6);
    if (lane == 0 && n_slow_words) atomicAdd(&status[kStatusSlowWords], n_slow_words);
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
template <class Enc>
__global__ void __launch_bounds__(kThreads, kEmitCtasPerSm)
encode_emit_kernel(Enc enc, const uint8_t *__restrict__ arena, const uint32_t *__restrict__ word_off, uint32_t n_words,
                   uint32_t *__restrict__ out_ids, uint64_t out_cap, uint32_t *__restrict__ out_tok_off, uint32_t tok_base,
                   EncodeWorkspace ws) {
    __shared__ __align__(16) uint32_t s_compact[kWarps][kCompactTokens + 4];   // per warp: the tile's ids in output order
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * kThreads + threadIdx.x) >> 5, n_warps = (gridDim.x * kThreads) >> 5;
    const bool tok_off_vec = out_tok_off && (((uintptr_t)out_tok_off & 7) == 0);
    uint32_t *compact = s_compact[threadIdx.x >> 5];

    // the per-word records and the tile's output position are fetched one tile ahead
    auto load_tile = [&](uint32_t t, uint32_t &p0, uint32_t &p1, uint64_t &b) {
        const uint32_t w0 = t * kTileWords, tw = min((uint32_t)kTileWords, n_words - w0), i = lane * kWordsPerThread;
        p0 = p1 = 0u;
        if (i + 1 < tw) { const uint2 p = *reinterpret_cast<const uint2 *>(ws.packed + w0 + i); p0 = p.x; p1 = p.y; }
        else if (i < tw) p0 = ws.packed[w0 + i];
        b = ws.group_base[t / kGroupTiles] + ws.tile_total[t];
    };
    uint32_t pp0 = 0, pp1 = 0; uint64_t pbase = 0;
    if (warp_global < ws.n_tiles) load_tile(warp_global, pp0, pp1, pbase);
    for (uint32_t tile = warp_global; tile < ws.n_tiles; tile += n_warps) {
        const uint32_t w_tile = tile * kTileWords;
        const uint32_t tile_words = min((uint32_t)kTileWords, n_words - w_tile);
        const uint32_t i0 = lane * kWordsPerThread;
        const uint32_t packed[kWordsPerThread] = {pp0, pp1};
        const uint64_t base = pbase;
        if (tile + n_warps < ws.n_tiles) load_tile(tile + n_warps, pp0, pp1, pbase);
        uint32_t kind[kWordsPerThread], ntok[kWordsPerThread], arg[kWordsPerThread];
        uint4 ra[kWordsPerThread], rb[kWordsPerThread];
#pragma unroll
        for (int j = 0; j < kWordsPerThread; ++j) {
            kind[j] = packed[j] >> 29;
            arg[j] = packed[j] & 0x1FFFFFFFu;
            ntok[j] = (kind[j] == kWordHit || kind[j] == kWordHit16 || kind[j] == kWordRecompute) ? (arg[j] >> 23)
                      : (kind[j] == kWordLong && !Enc::kScratchLong) ? arg[j] : 0u;
            ra[j] = rb[j] = make_uint4(0, 0, 0, 0);
            const MemoEntry *e = ws.memo + (arg[j] & kMemoSlotMask);
            if (kind[j] == kWordHit16) {                            // one or two scattered loads; both words' loads in flight together
                if (ntok[j] > 0) ra[j] = ld_ca_u32x4(&e->ids16[0]);
                if (ntok[j] > 8) rb[j] = ld_ca_u32x4(&e->ids16[8]);
            } else if (kind[j] == kWordHit) {
                if (ntok[j] > 0) ra[j] = ld_ca_u32x4(&e->tok[0]);
                if (ntok[j] > 4) rb[j] = ld_ca_u32x4(&e->tok[4]);
            }
            if constexpr (Enc::kScratchLong) {
                if (kind[j] == kWordLong || kind[j] == kWordLongB) ntok[j] = ws.long_scratch[(unsigned long long)arg[j] << 4];
            }
        }
        uint32_t count = 0, run[kWordsPerThread];
#pragma unroll
        for (int j = 0; j < kWordsPerThread; ++j) { run[j] = count; count += ntok[j]; }
        uint32_t incl = count;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += v; }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31), excl = incl - count;
#pragma unroll
        for (int j = 0; j < kWordsPerThread; ++j) run[j] += excl;
        const bool fits_out = base + total <= out_cap;
        const bool use_compact = total <= (uint32_t)kCompactTokens;
        if (out_tok_off) {
            const uint32_t o0 = tok_base + (uint32_t)base;
            if (tok_off_vec && i0 + kWordsPerThread <= tile_words)
                *reinterpret_cast<uint2 *>(out_tok_off + w_tile + i0) = make_uint2(o0 + run[0], o0 + run[1]);
            else {
#pragma unroll
                for (int j = 0; j < kWordsPerThread; ++j) if (i0 + j < tile_words) out_tok_off[w_tile + i0 + j] = o0 + run[j];
            }
        }
        if (!fits_out) continue;                                                    // warp-uniform (status set by the scan)
        // the tile's ids are laid out in shared memory with the 16-byte phase of their destination, so that the copy
        // out is LDS.128 -> STG.128 without bank conflicts
        const uint32_t sh = (uint32_t)((uintptr_t)(out_ids + base) >> 2) & 3u;
        uint32_t *cdst = compact + sh;
#pragma unroll
        for (int j = 0; j < kWordsPerThread; ++j) {
            if (kind[j] == kWordHit16) {
                // two copies of the same code so that the common case compiles to shared-memory stores (STS)
                if (use_compact) store_hit16_ids<true>(cdst + run[j], ntok[j], ra[j], rb[j]);
                else store_hit16_ids<false>(out_ids + base + run[j], ntok[j], ra[j], rb[j]);
            } else if (kind[j] == kWordHit) {
                if (use_compact) store_hit_ids(cdst + run[j], ntok[j], ra[j], rb[j], ws.memo + (arg[j] & kMemoSlotMask));
                else store_hit_ids(out_ids + base + run[j], ntok[j], ra[j], rb[j], ws.memo + (arg[j] & kMemoSlotMask));
            } else if (kind[j] == kWordRecompute || (!Enc::kScratchLong && kind[j] == kWordLong)) {
                uint32_t *dst = use_compact ? cdst + run[j] : out_ids + base + run[j];
                const uint32_t b0 = __ldg(word_off + w_tile + i0 + j), b1 = __ldg(word_off + w_tile + i0 + j + 1);
                emit_slow(enc, arena + b0, b1 - b0, kind[j], ntok[j], dst);
            }
        }
        if constexpr (Enc::kScratchLong) {                                          // long results: the warp copies them together
#pragma unroll
            for (int j = 0; j < kWordsPerThread; ++j) {
                uint32_t lm = __ballot_sync(0xffffffffu, kind[j] == kWordLong || kind[j] == kWordLongB);
                while (lm) {
                    const uint32_t owner = __ffs(lm) - 1; lm &= lm - 1;
                    const uint32_t ln = __shfl_sync(0xffffffffu, ntok[j], owner), lrun = __shfl_sync(0xffffffffu, run[j], owner);
                    const uint32_t larg = __shfl_sync(0xffffffffu, arg[j], owner), lkind = __shfl_sync(0xffffffffu, kind[j], owner);
                    const uint32_t lb0 = __ldg(word_off + w_tile + owner * kWordsPerThread + j);
                    const uint32_t lnb = __ldg(word_off + w_tile + owner * kWordsPerThread + j + 1) - lb0;
                    const uint32_t *src = ws.long_scratch + ((unsigned long long)larg << 4) + kLongHeader + (lkind == kWordLongB ? lnb : 0u);
                    uint32_t *dst = use_compact ? cdst + lrun : out_ids + base + lrun;
                    for (uint32_t k = lane; k < ln; k += 32) dst[k] = src[k];
                }
            }
        }
        __syncwarp();
        if (use_compact) {
            uint32_t *dsta = out_ids + base - sh;                                   // 16-byte aligned
            const uint32_t end = sh + total, nvec = (end + 3) >> 2;
            for (uint32_t v = lane; v < nvec; v += 32) {
                const uint32_t c = 4 * v;
                const uint4 q = *reinterpret_cast<const uint4 *>(compact + c);
                if (c >= sh && c + 4 <= end) *reinterpret_cast<uint4 *>(dsta + c) = q;
                else {                                                              // first / last vector of the tile
                    if (c >= sh && c < end) dsta[c] = q.x;
                    if (c + 1 >= sh && c + 1 < end) dsta[c + 1] = q.y;
                    if (c + 2 >= sh && c + 2 < end) dsta[c + 2] = q.z;
                    if (c + 3 < end) dsta[c + 3] = q.w;
                }
            }
        }
        __syncwarp();                                            // the compact buffer is reused by the next tile
    }
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
    // SWT_TIMING=1: per-kernel CUDA-event times of this call on stderr (diagnostic; synchronises)
    static const bool timing = getenv("SWT_TIMING") != nullptr;
    cudaEvent_t ev[6];
    if (timing) for (auto &e : ev) cudaEventCreate(&e);
    if (timing) cudaEventRecord(ev[0], st);
    memo_clear_kernel<<<kNumSMs * 8, 256, 0, st>>>(ws.memo, ws.memo_mask + 1, ws.long_cursor);
    if (timing) cudaEventRecord(ev[1], st);
    encode_count_kernel<Enc><<<(int)std::min<uint32_t>((uint32_t)grid1, n_ctas), kThreads, 0, st>>>(enc, d_arena, d_word_off, n_words, ws, d_status);
    if (timing) cudaEventRecord(ev[2], st);
    encode_scan_groups_kernel<<<n_groups, 256, 0, st>>>(ws);
    if (timing) cudaEventRecord(ev[3], st);
    encode_scan_top_kernel<<<1, 1024, 0, st>>>(ws, n_words, d_out_tok_off, tok_base, out_cap, d_status);
    if (timing) cudaEventRecord(ev[4], st);
    encode_emit_kernel<Enc><<<(int)std::min<uint32_t>((uint32_t)grid2, n_ctas), kThreads, 0, st>>>(enc, d_arena, d_word_off, n_words, d_out_ids, out_cap,
                                                                                                   d_out_tok_off, tok_base, ws);
    if (timing) {
        cudaEventRecord(ev[5], st); cudaEventSynchronize(ev[5]);
        float t[5];
        for (int i = 0; i < 5; ++i) cudaEventElapsedTime(&t[i], ev[i], ev[i + 1]);
        fprintf(stderr, "[swt timing] clear %.3f count %.3f scan %.3f+%.3f emit %.3f ms (grid %d/%d, %u tiles)\n", t[0], t[1], t[2], t[3], t[4], grid1, grid2, ws.n_tiles);
        for (auto &e : ev) cudaEventDestroy(e);
    }
    SWT_CUDA_OK(cudaGetLastError());
    return SWT_OK;
}
#endif  // __CUDACC__

}  // namespace swt
