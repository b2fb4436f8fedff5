// encode.cuh -- the tile kernel shared by the two tokenize paths (HP-1 FastBPE, HP-2 FastWP).
//
// Data flow of one launch (north-star subsystem 1: packed word-offset/byte arena):
//
//   arena bytes + u32 word offsets --(tile = 1024 consecutive words per CTA, 2 consecutive words per thread)-->
//   phase A  per word: look the word up in the word-type memo; on a miss encode it (rank table / trie walk)
//            and publish the ids; only the token COUNT is kept
//   scan     CTA exclusive scan of the per-thread counts, then a warp-parallel decoupled look-back over a
//            64-bit tile-state array gives the tile's global token offset (single pass, no second kernel)
//   phase B  per word: copy the ids (memo entry -> shared memory) at their tile-local position
//   store    the tile's ids leave shared memory with fully coalesced writes; u32 token offsets go out as 16 B stores
//
// The grid is persistent (kNumSMs x resident CTAs); tiles are handed out by an atomic ticket, so the look-back
// never waits on a tile that has not started.
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

#include "common.cuh"

namespace swt {

constexpr int kThreads = 512;
constexpr int kWordsPerThread = 2;
constexpr int kTileWords = kThreads * kWordsPerThread;   // 1024 words per tile
constexpr int kShortBytes = 32;        // words up to this many bytes are encoded by one thread (and memoised)
constexpr int kCompactTokens = 8192;   // tile token totals up to this are assembled in smem before the store
constexpr int kMemoTokens = 54;        // >= kShortBytes: every word of up to 32 bytes fits (ids <= bytes)
constexpr int kMemoProbes = 8;
constexpr int kSlowCap = 512;          // per tile: words resolved by the dense slow pass (more are resolved inline)

// status words written by the encode kernels
enum { kStatusCode = 0, kStatusTokens = 1, kStatusH6 = 2, kStatusTokensHi = 3, kStatusMemoTypes = 4 };

struct alignas(256) MemoEntry {         // 256 bytes; a typical hit touches the first 32-96 bytes only
    unsigned long long lo, hi;          // CAS key: word bytes 0..7 | bytes 8..14 + (length << 56); 0/0 == empty
    uint32_t meta;                      // 0 = claimed, not published; else (n_tokens + 1) | (h6 << 8); ~0 = not cacheable
    uint32_t tail_last;                 // byte 31
    uint32_t tok01[2];                  // ids 0 and 1: a word of up to two ids is served by the first 32-byte sector
    unsigned long long tail_a, tail_b;  // bytes 15..22 | 23..30 (words longer than 15 bytes; verified after the key)
    uint32_t tok[kMemoTokens - 2];      // ids 2.. at byte 48
};
struct MemoKey { unsigned long long lo, hi, tail_a, tail_b; uint32_t tail_last; uint32_t nbytes; };

struct EncodeWorkspace {
    uint64_t *tile_state;            // n_tiles
    uint32_t *ticket;                // 1
    unsigned long long *long_cursor; // 1 (BPE: allocation cursor into long_scratch, in u32 units)
    MemoEntry *memo; uint32_t memo_mask;   // memo_mask == 0: memo disabled
    uint32_t *long_scratch;          // 2 x long bytes (BPE symbol ping-pong buffers for long words)
    uint64_t long_scratch_elems;
    uint32_t n_tiles;
    size_t zero_bytes;               // prefix of the workspace that must be zeroed before a launch
};

size_t encode_workspace_layout(uint32_t n_words, uint64_t long_bytes, void *base, EncodeWorkspace *ws);
int encode_grid(const void *kernel, int block, size_t dyn_smem);

#ifdef __CUDACC__
enum { kMemoHit = 0, kMemoClaimed = 1, kMemoMiss = 2 };

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
__device__ __forceinline__ uint4 ld_cg_u32x4(const void *p) {
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
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
// meta is read with a STRONG RELAXED load served at L2, not an acquire: ld.acquire.gpu compiles to LDG + CCTL.IVALL
// (a full L1 invalidate per probe).  Ordering comes from the writer's release (ids are performed at L2 before meta)
// plus the reader's control dependency (ids are loaded, at L2, only after a valid meta has been observed).
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

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
    k.hi = (w1 & ((1ull << 56) - 1)) | ((unsigned long long)nbytes << 56);
    k.tail_a = (w1 >> 56) | (w2 << 8);
    k.tail_b = (w2 >> 56) | (w3 << 8);
    k.tail_last = (uint32_t)(w3 >> 56);
    k.nbytes = nbytes;
}

// Probes the memo.  kMemoHit: slot/meta describe a published entry for exactly this word, t01 holds its first two
// ids.  kMemoClaimed: this thread now owns `slot` and must call memo_publish after encoding.  kMemoMiss: encode
// directly, publish nothing.  `first` is the probe index to start from (the caller may have checked probe 0 itself).
static __device__ __noinline__ int memo_probe(const EncodeWorkspace &ws, const MemoKey &key, uint32_t &slot, uint32_t &meta_out, uint2 &t01) {
    uint32_t h = (uint32_t)mix64(key.lo ^ (key.hi * 0x9E3779B97F4A7C15ull)) & ws.memo_mask;
    for (int probe = 0; probe < kMemoProbes; ++probe, h = (h + 1) & ws.memo_mask) {
        MemoEntry *e = ws.memo + h;
        unsigned long long klo, khi;
        ld_cg_u64x2(e, klo, khi);
        uint4 m = ld_cg_u32x4(&e->meta);                                 // meta, tail_last, id0, id1 (same sector as the key)
        if (klo == 0 && khi == 0) {
            cas128(e, key.lo, key.hi, klo, khi);
            if (klo == 0 && khi == 0) { slot = h; return kMemoClaimed; }
            m.x = 0;                                                     // lost the race: the winner has not published yet
        }
        if (klo != key.lo || khi != key.hi) continue;
        if (m.x == 0 || m.x == 0xFFFFFFFFu) return kMemoMiss;            // not published yet / not cacheable
        if (key.nbytes > 15) {                                           // same 15-byte prefix and length: check the rest
            unsigned long long ta, tb;
            ld_cg_u64x2(&e->tail_a, ta, tb);
            if (ta != key.tail_a || tb != key.tail_b || m.y != key.tail_last) continue;
        }
        slot = h; meta_out = m.x; t01 = make_uint2(m.z, m.w);
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
    e->tok01[0] = ntok > 0 ? buf[0] : 0u; e->tok01[1] = ntok > 1 ? buf[1] : 0u;
    for (uint32_t k = 2; k < ntok; ++k) e->tok[k - 2] = buf[k];
    st_release_u32(&e->meta, (ntok + 1) | (h6 << 8));
    atomicAdd(&status[kStatusMemoTypes], 1u);
    return true;
}

// ---- warp-parallel decoupled look-back: all 32 lanes of one warp call this -----------------------------------------
// The aggregate is published first; later one warp inspects 32 predecessors per step until a tile with a published
// inclusive prefix is found and returns (in every lane) the exclusive prefix of `tile`.
// step 1 (one thread, right after the CTA scan): make the tile's aggregate visible to later tiles at once
__device__ __forceinline__ void tile_publish_aggregate(uint64_t *tile_state, uint32_t tile, uint64_t aggregate) {
    st_relaxed_u64(&tile_state[tile], ((tile == 0 ? kTilePrefix : kTileAggregate) << 62) | aggregate);
}
// step 2 (all 32 lanes of one warp, any time later): look back, publish the inclusive prefix, return the exclusive one
__device__ __forceinline__ uint64_t tile_prefix_warp(uint64_t *tile_state, uint32_t tile, uint64_t aggregate, uint32_t *err) {
    const uint32_t lane = threadIdx.x & 31;
    if (tile == 0) return 0;
    uint64_t running = 0;
    int64_t p = (int64_t)tile - 1;                  // lane l looks at tile p - l
    uint32_t spins = 0;
    for (;;) {
        const int64_t idx = p - (int64_t)lane;
        const uint64_t s = idx >= 0 ? ld_relaxed_u64(&tile_state[idx]) : (kTilePrefix << 62);   // before tile 0: prefix 0
        const uint64_t st = s >> 62;
        const uint32_t inv = __ballot_sync(0xffffffffu, st == kTileInvalid);
        const uint32_t pre = __ballot_sync(0xffffffffu, st == kTilePrefix);
        const uint32_t first_pre = pre ? (uint32_t)__ffs(pre) - 1 : 32u;
        const uint32_t first_inv = inv ? (uint32_t)__ffs(inv) - 1 : 32u;
        if (first_inv < first_pre) {                                        // a needed predecessor is not there yet
            if (++spins > (1u << 24)) { if (lane == 0) atomicExch(err, (uint32_t)SWT_ERR_INTERNAL); break; }   // never hang
            __nanosleep(40);
            continue;
        }
        uint64_t v = lane <= first_pre ? (s & kTileValueMask) : 0ull;      // aggregates up to and including the prefix tile
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        running += v;
        if (first_pre < 32u) break;
        p -= 32;
    }
    if (lane == 0) st_relaxed_u64(&tile_state[tile], (kTilePrefix << 62) | ((running + aggregate) & kTileValueMask));
    return running;
}

// per-word bookkeeping between phase A and phase B
enum : uint32_t { kWordNone = 0u, kWordHit = 1u, kWordRecompute = 2u, kWordLong = 3u, kWordLongB = 4u, kWordPending = 5u };

struct TileSmem {                         // dynamic shared memory of the tile kernel (37 KB)
    uint32_t compact[kCompactTokens];     // the tile's ids in output order
    uint32_t off[kTileWords + 4];         // the tile's word offsets
    uint32_t scan[36];
    uint32_t misc[8];
    uint64_t base;
    uint32_t n_slow;                      // words of this tile that missed the fast path ...
    uint32_t slow_list[kSlowCap];         // ... their tile-local index ...
    uint32_t slow_res[kSlowCap][4];       // ... and, once resolved by a dense pass, {kind | ntok << 8, slot, id0, id1}
};

// phase A slow path (first probe did not hit): full probe, then direct encode + publish.  Kept out of line so
// that the 4x unrolled fast path stays small.
struct SlowResult { uint32_t kind, ntok, slot, h6; uint2 t01; };
template <class Enc>
__device__ __noinline__ SlowResult resolve_slow(const Enc &enc, const EncodeWorkspace &ws, const uint8_t *arena, uint32_t b0,
                                                uint32_t nbytes, uint32_t arena_end, uint32_t *buf, uint32_t *status) {
    SlowResult r; r.kind = kWordRecompute; r.ntok = 0; r.slot = 0; r.h6 = 0; r.t01 = make_uint2(0, 0);
    int m = kMemoMiss; uint32_t meta = 0; MemoKey key;
    if (ws.memo_mask && nbytes >= 1) { memo_key(arena, b0, nbytes, arena_end, key); m = memo_probe(ws, key, r.slot, meta, r.t01); }
    if (m == kMemoHit) { r.kind = kWordHit; r.ntok = (meta & 0xFFu) - 1; r.h6 = meta >> 8; return r; }
    r.ntok = enc.encode_short(arena + b0, nbytes, buf, r.h6);
    if (m == kMemoClaimed && memo_publish(ws, r.slot, key, buf, r.ntok, r.h6, status)) {
        r.kind = kWordHit; r.t01 = make_uint2(r.ntok > 0 ? buf[0] : 0u, r.ntok > 1 ? buf[1] : 0u);
    }
    return r;
}
// phase B slow paths: re-encode a word whose ids were not kept, or emit a long word
template <class Enc>
__device__ __noinline__ uint32_t emit_slow(const Enc &enc, const EncodeWorkspace &ws, const uint8_t *word, uint32_t nbytes,
                                           uint32_t kind, uint32_t ntok, uint32_t slot, uint32_t *buf, uint32_t *dst) {
    uint32_t h6 = 0;
    if (kind == kWordRecompute) {
        uint32_t dummy = 0;
        const uint32_t n = enc.encode_short(word, nbytes, buf, dummy);
        for (uint32_t k = 0; k < n; ++k) dst[k] = buf[k];
    } else if constexpr (!Enc::kCoopLong) {
        enc.long_emit(word, nbytes, dst, ntok, h6);
    } else {
        const uint32_t *src = ws.long_scratch + ((unsigned long long)slot << 1) + (kind == kWordLongB ? nbytes : 0u);
        for (uint32_t k = 0; k < ntok; ++k) dst[k] = src[k];
    }
    return h6;
}

// ids of a memo hit -> dst (ids 0-1 arrived with the probe, 2-9 were prefetched, the rest is fetched here)
__device__ __forceinline__ void store_hit_ids(uint32_t *dst, uint32_t n, uint2 t01, uint4 c0, uint4 c1, const MemoEntry *e) {
    if (n > 0) dst[0] = t01.x;
    if (n > 1) dst[1] = t01.y;
    if (n > 2) dst[2] = c0.x;
    if (n > 3) dst[3] = c0.y;
    if (n > 4) dst[4] = c0.z;
    if (n > 5) dst[5] = c0.w;
    if (n > 6) dst[6] = c1.x;
    if (n > 7) dst[7] = c1.y;
    if (n > 8) dst[8] = c1.z;
    if (n > 9) dst[9] = c1.w;
    for (uint32_t k0 = 8; k0 + 2 < n; k0 += 4) {
        const uint4 v = ld_ca_u32x4(&e->tok[k0]);
        dst[k0 + 2] = v.x;
        if (k0 + 3 < n) dst[k0 + 3] = v.y;
        if (k0 + 4 < n) dst[k0 + 4] = v.z;
        if (k0 + 5 < n) dst[k0 + 5] = v.w;
    }
}

// ---- the tile kernel ---------------------------------------------------------------------------------------------------
// Enc provides
//   uint32_t encode_short(const uint8_t *p, uint32_t nbytes, uint32_t *buf /*thread-local, kShortBytes*/, uint32_t &h6) const
//   static constexpr bool kCoopLong
//   kCoopLong == false:  uint32_t long_count(p, nbytes) const;  void long_emit(p, nbytes, uint32_t *dst, uint32_t cap, uint32_t &h6) const
//   kCoopLong == true :  uint32_t encode_long_coop(p, nbytes, bufA, bufB, uint32_t **result, uint32_t *sh_scan, uint32_t *sh_misc) const
template <class Enc>
__global__ void __launch_bounds__(kThreads, 2)
encode_tiles_kernel(Enc enc, const uint8_t *__restrict__ arena, const uint32_t *__restrict__ word_off, uint32_t n_words,
                    uint32_t *__restrict__ out_ids, uint64_t out_cap, uint32_t *__restrict__ out_tok_off, uint32_t tok_base,
                    EncodeWorkspace ws, uint32_t *status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileSmem &sm = *reinterpret_cast<TileSmem *>(smem_raw);
    const uint32_t tid = threadIdx.x;
    const uint32_t arena_end = word_off[n_words];
    const bool tok_off_vec = out_tok_off && (((uintptr_t)out_tok_off & 15) == 0);
    const bool use_memo = ws.memo_mask != 0;
    uint32_t h6 = 0;
    uint32_t buf[kShortBytes];                              // thread-local scratch for one directly encoded word

    // Tiles are handed out by an atomic ticket: a tile only starts once a CTA is free to run it, so every predecessor
    // of a running tile is itself running or finished and the look-back cannot deadlock.
    // Tiles are handed out by an atomic ticket: a tile only starts once a CTA is free to run it, so every predecessor
    // of a running tile is itself running or finished and the look-back cannot deadlock.
    uint32_t next_ticket = 0;
    if (tid == 0) next_ticket = atomicAdd(ws.ticket, 1u);
    for (;;) {
        if (tid == 0) { sm.misc[4] = next_ticket; sm.n_slow = 0; }
        __syncthreads();
        const uint32_t tile = sm.misc[4];
        if (tile >= ws.n_tiles) break;
        const uint32_t w_tile = tile * kTileWords;
        const uint32_t tile_words = min((uint32_t)kTileWords, n_words - w_tile);
        for (uint32_t i = tid; i <= tile_words; i += kThreads) sm.off[i] = word_off[w_tile + i];
        __syncthreads();

        // ---- phase A: counts.  Fast path = first memo probe hits; everything else goes through resolve_slow.
        uint32_t kind[kWordsPerThread], ntok[kWordsPerThread], slot[kWordsPerThread];
        uint2 t01[kWordsPerThread];
        unsigned long long klo[kWordsPerThread], khi[kWordsPerThread];      // word key, then (after the xor) key difference
        uint32_t nb[kWordsPerThread], b0s[kWordsPerThread];
        uint4 mt[kWordsPerThread];
        uint32_t count = 0; bool has_long = false;
#pragma unroll
        for (int j = 0; j < kWordsPerThread; ++j) {
            const uint32_t i = tid * kWordsPerThread + j;
            kind[j] = kWordNone; ntok[j] = 0; slot[j] = 0; t01[j] = make_uint2(0, 0);
            b0s[j] = i < tile_words ? sm.off[i] : 0u;
            nb[j] = i < tile_words ? sm.off[i + 1] - b0s[j] : 0xFFFFFFFFu;          // 0xFFFFFFFF: no word
            klo[j] = khi[j] = ~0ull; mt[j] = make_uint4(0, 0, 0, 0);
        }
        if (use_memo) {
            // words of 1..15 bytes: request the (up to three) aligned 8-byte words of all four words, then the first memo
            // probe of all four, before looking at any result
            unsigned long long r0[kWordsPerThread], r1[kWordsPerThread], r2[kWordsPerThread];
            bool fastj[kWordsPerThread];
#pragma unroll
            for (int j = 0; j < kWordsPerThread; ++j) {
                fastj[j] = nb[j] >= 1 && nb[j] <= 15 && (uint64_t)b0s[j] + 24 <= arena_end;
                r0[j] = r1[j] = r2[j] = 0;
                if (fastj[j]) {
                    const uintptr_t a = (uintptr_t)(arena + b0s[j]);
                    const unsigned long long *q = (const unsigned long long *)(a & ~(uintptr_t)7);
                    r0[j] = __ldg(q); r1[j] = __ldg(q + 1);
                    if (nb[j] + (uint32_t)(a & 7) > 16) r2[j] = __ldg(q + 2);
                }
            }
#pragma unroll
            for (int j = 0; j < kWordsPerThread; ++j) {
                if (fastj[j]) {
                    const uint32_t sh = (uint32_t)((uintptr_t)(arena + b0s[j]) & 7) * 8;
                    unsigned long long lo = r0[j], hi = r1[j], elo, ehi;
                    if (sh) { lo = (r0[j] >> sh) | (r1[j] << (64 - sh)); hi = (r1[j] >> sh) | (r2[j] << (64 - sh)); }
                    if (nb[j] < 8) { lo &= (1ull << (8 * nb[j])) - 1; hi = 0; }
                    else hi = nb[j] == 8 ? 0ull : hi & ((1ull << (8 * (nb[j] - 8))) - 1);
                    hi |= (unsigned long long)nb[j] << 56;
                    slot[j] = (uint32_t)mix64(lo ^ (hi * 0x9E3779B97F4A7C15ull)) & ws.memo_mask;
                    const MemoEntry *e = ws.memo + slot[j];
                    ld_ca_u64x2(e, elo, ehi);
                    mt[j] = ld_ca_u32x4(&e->meta);
                    klo[j] = lo ^ elo; khi[j] = hi ^ ehi;                            // zero iff the entry holds this word
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kWordsPerThread; ++j) {
            if (nb[j] == 0xFFFFFFFFu) continue;
            if (nb[j] > (uint32_t)kShortBytes) {
                kind[j] = kWordLong; has_long = true;
                if constexpr (!Enc::kCoopLong) ntok[j] = enc.long_count(arena + b0s[j], nb[j]);
            } else if ((klo[j] | khi[j]) == 0 && mt[j].x != 0 && mt[j].x != 0xFFFFFFFFu) {
                kind[j] = kWordHit; ntok[j] = (mt[j].x & 0xFFu) - 1; h6 += mt[j].x >> 8; t01[j] = make_uint2(mt[j].z, mt[j].w);
            } else {
                // not served by the first probe (longer than 15 bytes, hash collision, first occurrence): defer to the
                // dense pass below so that these few words do not serialise whole warps one lane at a time
                const uint32_t q = atomicAdd(&sm.n_slow, 1u);
                if (q < (uint32_t)kSlowCap) { sm.slow_list[q] = tid * kWordsPerThread + j; kind[j] = kWordPending; slot[j] = q; }
                else {
                    const SlowResult r = resolve_slow(enc, ws, arena, b0s[j], nb[j], arena_end, buf, status);
                    kind[j] = r.kind; ntok[j] = r.ntok; slot[j] = r.slot; t01[j] = r.t01; h6 += r.h6;
                }
            }
        }
        __syncthreads();
        {
            const uint32_t n_slow = min(sm.n_slow, (uint32_t)kSlowCap);
            for (uint32_t q = tid; q < n_slow; q += kThreads) {                      // one deferred word per thread, lanes dense
                const uint32_t i = sm.slow_list[q];
                const uint32_t b0 = sm.off[i];
                const SlowResult r = resolve_slow(enc, ws, arena, b0, sm.off[i + 1] - b0, arena_end, buf, status);
                h6 += r.h6;
                sm.slow_res[q][0] = r.kind | (r.ntok << 8); sm.slow_res[q][1] = r.slot; sm.slow_res[q][2] = r.t01.x; sm.slow_res[q][3] = r.t01.y;
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kWordsPerThread; ++j) {
            if (kind[j] == kWordPending) {
                const uint32_t q = slot[j];
                const uint32_t kn = sm.slow_res[q][0];
                kind[j] = kn & 0xFFu; ntok[j] = kn >> 8; slot[j] = sm.slow_res[q][1]; t01[j] = make_uint2(sm.slow_res[q][2], sm.slow_res[q][3]);
            }
            count += ntok[j];
        }
        if constexpr (Enc::kCoopLong) {
            // long words: the whole CTA works on one word at a time in global scratch (rare)
            if (__syncthreads_or(has_long)) {
                for (uint32_t i = 0; i < tile_words; ++i) {
                    const uint32_t b0 = sm.off[i], nbytes = sm.off[i + 1] - b0;
                    if (nbytes <= (uint32_t)kShortBytes) continue;                      // CTA-uniform
                    if (tid == 0) {
                        const unsigned long long so = atomicAdd(ws.long_cursor, 2ull * nbytes);
                        sm.misc[6] = (uint32_t)(so >> 1); sm.misc[7] = (so + 2ull * nbytes <= ws.long_scratch_elems) ? 1u : 0u;
                    }
                    __syncthreads();
                    const unsigned long long so = (unsigned long long)sm.misc[6] << 1;
                    const bool fits = sm.misc[7] != 0;
                    __syncthreads();
                    uint32_t *res = nullptr; uint32_t c = 0;
                    if (fits) c = enc.encode_long_coop(arena + b0, nbytes, ws.long_scratch + so, ws.long_scratch + so + nbytes, &res, sm.scan, sm.misc);
                    else if (tid == 0) atomicExch(&status[kStatusCode], (uint32_t)SWT_ERR_CAPACITY);
                    if (tid == i / kWordsPerThread) {
#pragma unroll
                        for (int j = 0; j < kWordsPerThread; ++j) if ((uint32_t)j == i % kWordsPerThread) {
                            ntok[j] = c; count += c; slot[j] = (uint32_t)(so >> 1);
                            kind[j] = (res == ws.long_scratch + so) ? kWordLong : kWordLongB;
                            if (!fits) kind[j] = kWordNone;
                        }
                    }
                    __syncthreads();
                }
            }
        }

        // ---- scan.  The look-back (warp 0) mostly WAITS for earlier tiles, so it runs after warp 0 has assembled its
        // own share of the tile's ids in shared memory; only an oversized tile (rare) needs the base first.
        uint32_t total, excl = block_exclusive_scan(count, sm.scan, &total);
        if (tid == 0) tile_publish_aggregate(ws.tile_state, tile, total);
        const bool use_compact = total <= (uint32_t)kCompactTokens;
        uint64_t base = 0; bool fits_out = true;
        if (!use_compact) {
            if (tid < 32) { const uint64_t b = tile_prefix_warp(ws.tile_state, tile, total, &status[kStatusCode]); if (tid == 0) sm.base = b; }
            __syncthreads();
            base = sm.base; fits_out = base + total <= out_cap;
        }

        // ---- phase B: ids to their tile-local position.  Ids 0-1 came with the probe; ids 2-9 of all four words are
        // requested before any is stored; the rest (rare) in a loop.
        uint32_t run[kWordsPerThread];
        {
            uint32_t r = excl;
#pragma unroll
            for (int j = 0; j < kWordsPerThread; ++j) { run[j] = r; r += ntok[j]; }
        }
        if (fits_out) {
            uint4 c0[kWordsPerThread], c1[kWordsPerThread];
#pragma unroll
            for (int j = 0; j < kWordsPerThread; ++j) {
                c0[j] = c1[j] = make_uint4(0, 0, 0, 0);
                if (kind[j] == kWordHit && ntok[j] > 2) {
                    const MemoEntry *e = ws.memo + slot[j];
                    c0[j] = ld_ca_u32x4(&e->tok[0]);
                    if (ntok[j] > 6) c1[j] = ld_ca_u32x4(&e->tok[4]);
                }
            }
#pragma unroll
            for (int j = 0; j < kWordsPerThread; ++j) {
                if (kind[j] == kWordNone) continue;
                if (kind[j] == kWordHit) {
                    // two copies of the same code so that the common case compiles to shared-memory stores (STS)
                    if (use_compact) store_hit_ids(sm.compact + run[j], ntok[j], t01[j], c0[j], c1[j], ws.memo + slot[j]);
                    else store_hit_ids(out_ids + base + run[j], ntok[j], t01[j], c0[j], c1[j], ws.memo + slot[j]);
                } else {
                    uint32_t *dst = use_compact ? sm.compact + run[j] : out_ids + base + run[j];
                    h6 += emit_slow(enc, ws, arena + b0s[j], nb[j], kind[j], ntok[j], slot[j], buf, dst);
                }
            }
        }
        if (use_compact && tid < 32) { const uint64_t b = tile_prefix_warp(ws.tile_state, tile, total, &status[kStatusCode]); if (tid == 0) sm.base = b; }
        __syncthreads();
        if (use_compact) { base = sm.base; fits_out = base + total <= out_cap; }
        if (!fits_out && tid == 0) atomicExch(&status[kStatusCode], (uint32_t)SWT_ERR_CAPACITY);
        if (out_tok_off) {
            const uint32_t i0 = tid * kWordsPerThread;
            const uint32_t o0 = tok_base + (uint32_t)base;
            if (tok_off_vec && i0 + kWordsPerThread <= tile_words) {
                if constexpr (kWordsPerThread == 4)
                    *reinterpret_cast<uint4 *>(out_tok_off + w_tile + i0) = make_uint4(o0 + run[0], o0 + run[1], o0 + run[2 % kWordsPerThread], o0 + run[3 % kWordsPerThread]);
                else
                    *reinterpret_cast<uint2 *>(out_tok_off + w_tile + i0) = make_uint2(o0 + run[0], o0 + run[1]);
            } else {
#pragma unroll
                for (int j = 0; j < kWordsPerThread; ++j) if (i0 + j < tile_words) out_tok_off[w_tile + i0 + j] = o0 + run[j];
            }
        }
        // take the next ticket now so that its latency hides behind the store of this tile.  (It must not be taken any
        // earlier: a tile that holds a ticket without running delays the look-back of every later tile -- measured.)
        if (tid == 0) next_ticket = atomicAdd(ws.ticket, 1u);
        if (use_compact && fits_out) {
            // coalesced store: scalar head up to 16-byte alignment of the destination, then 128-bit stores
            uint32_t *dst = out_ids + base;
            const uint32_t head = min(total, (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15) >> 2);
            if (tid < head) dst[tid] = sm.compact[tid];
            const uint32_t nvec = (total - head) >> 2;
            for (uint32_t v = tid; v < nvec; v += kThreads) {
                const uint32_t c = head + 4 * v;
                *reinterpret_cast<uint4 *>(dst + c) = make_uint4(sm.compact[c], sm.compact[c + 1], sm.compact[c + 2], sm.compact[c + 3]);
            }
            const uint32_t tail0 = head + 4 * nvec;
            if (tail0 + tid < total) dst[tail0 + tid] = sm.compact[tail0 + tid];
        }
        if (tile == ws.n_tiles - 1 && tid == 0) {
            const uint64_t grand = base + total;
            if (out_tok_off) out_tok_off[n_words] = tok_base + (uint32_t)grand;
            status[kStatusTokens] = (uint32_t)grand; status[kStatusTokensHi] = (uint32_t)(grand >> 32);
        }
        __syncthreads();
    }
    if (h6) atomicAdd(&status[kStatusH6], h6);
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
    SWT_CUDA_OK(cudaMemsetAsync(d_workspace, 0, ws.zero_bytes, st));
    SWT_CUDA_OK(cudaMemsetAsync(d_status, 0, 8 * sizeof(uint32_t), st));
    if (n_words == 0) {
        if (d_out_tok_off) SWT_CUDA_OK(cudaMemcpyAsync(d_out_tok_off, &tok_base, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        return SWT_OK;
    }
    static int grid = 0;
    if (!grid) {
        SWT_CUDA_OK(cudaFuncSetAttribute(encode_tiles_kernel<Enc>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem)));
        grid = encode_grid((const void *)encode_tiles_kernel<Enc>, kThreads, sizeof(TileSmem));
    }
    const int g = (int)std::min<uint64_t>((uint64_t)grid, ws.n_tiles);
    encode_tiles_kernel<Enc><<<g, kThreads, sizeof(TileSmem), st>>>(enc, d_arena, d_word_off, n_words, d_out_ids, out_cap, d_out_tok_off, tok_base, ws, d_status);
    SWT_CUDA_OK(cudaGetLastError());
    return SWT_OK;
}
#endif  // __CUDACC__

}  // namespace swt
