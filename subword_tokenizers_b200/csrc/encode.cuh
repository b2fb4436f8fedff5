// encode.cuh -- the tile kernel shared by the two tokenize paths (HP-1 FastBPE, HP-2 FastWP).
//
// Data flow of one launch (north-star subsystem 1: packed word-offset/byte arena):
//
//   arena bytes + u32 word offsets --(tile = 1024 consecutive words per CTA, 4 consecutive words per thread)-->
//   phase A  per word: look the word up in the word-type memo; on a miss encode it (rank table / trie walk)
//            and publish the ids; only the token COUNT is kept
//   scan     CTA exclusive scan of the per-thread counts, then a warp-parallel decoupled look-back over a
//            64-bit tile-state array gives the tile's global token offset (single pass, no second kernel)
//   phase B  per word: copy the ids (memo entry -> shared memory) at their tile-local position
//   store    the tile's ids leave shared memory with fully coalesced writes; u32 token offsets go out as 16 B stores
//
// Tiles are handed out by an atomic ticket so the grid is persistent
// (kNumSMs x resident CTAs) and the look-back never waits on a CTA that has not started.
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

constexpr int kThreads = 256;
constexpr int kWordsPerThread = 4;
constexpr int kTileWords = kThreads * kWordsPerThread;   // 1024 words per tile
constexpr int kShortBytes = 32;        // words up to this many bytes are encoded by one thread (and memoised)
constexpr int kCompactTokens = 8192;   // tile token totals up to this are assembled in smem before the store
constexpr int kMemoTokens = 52;        // >= kShortBytes: every word of up to 32 bytes fits (ids <= bytes)
constexpr int kMemoProbes = 8;

// status words written by the encode kernels
enum { kStatusCode = 0, kStatusTokens = 1, kStatusH6 = 2, kStatusTokensHi = 3, kStatusMemoTypes = 4 };

struct alignas(256) MemoEntry {         // 256 bytes; a typical hit touches the first 64-96 bytes only
    unsigned long long lo, hi;          // CAS key: word bytes 0..7 | bytes 8..14 + (length << 56); 0/0 == empty
    unsigned long long tail_a, tail_b;  // bytes 15..22 | 23..30 (words longer than 15 bytes; verified after the key)
    uint32_t meta;                      // 0 = claimed, not published; else (n_tokens + 1) | (h6 << 8); ~0 = not cacheable
    uint32_t tail_last;                 // byte 31
    uint32_t pad[2];
    uint32_t tok[kMemoTokens];          // at byte 48
};
static_assert(sizeof(MemoEntry) == 256, "MemoEntry must be 256 bytes");

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
int encode_grid(const void *kernel, int block);

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

// Probes the memo.  kMemoHit: slot/meta describe a published entry for exactly this word.  kMemoClaimed: this
// thread now owns `slot` and must call memo_publish after encoding.  kMemoMiss: encode directly, publish nothing.
__device__ __forceinline__ int memo_probe(const EncodeWorkspace &ws, const MemoKey &key, uint32_t &slot, uint32_t &meta_out) {
    uint32_t h = (uint32_t)mix64(key.lo ^ (key.hi * 0x9E3779B97F4A7C15ull)) & ws.memo_mask;
    for (int probe = 0; probe < kMemoProbes; ++probe, h = (h + 1) & ws.memo_mask) {
        MemoEntry *e = ws.memo + h;
        unsigned long long klo, khi;
        ld_cg_u64x2(e, klo, khi);
        uint32_t meta = ld_relaxed_u32(&e->meta);                        // issued together with the key load
        if (klo == 0 && khi == 0) {
            cas128(e, key.lo, key.hi, klo, khi);
            if (klo == 0 && khi == 0) { slot = h; return kMemoClaimed; }
            meta = 0;                                                    // lost the race: the winner has not published yet
        }
        if (klo != key.lo || khi != key.hi) continue;
        if (meta == 0 || meta == 0xFFFFFFFFu) return kMemoMiss;          // not published yet / not cacheable
        if (key.nbytes > 15) {                                           // same 15-byte prefix and length: check the rest
            unsigned long long ta, tb;
            ld_cg_u64x2(&e->tail_a, ta, tb);
            const uint4 m = ld_cg_u32x4(&e->meta);
            if (ta != key.tail_a || tb != key.tail_b || m.y != key.tail_last) continue;
        }
        slot = h; meta_out = meta;
        return kMemoHit;
    }
    return kMemoMiss;
}
// copies the ids of a published entry to dst[0..ntok)
__device__ __forceinline__ void memo_copy(const EncodeWorkspace &ws, uint32_t slot, uint32_t ntok, uint32_t *dst) {
    const MemoEntry *e = ws.memo + slot;
    for (uint32_t k0 = 0; k0 < ntok; k0 += 4) {
        const uint4 v = ld_cg_u32x4(&e->tok[k0]);
        dst[k0] = v.x;
        if (k0 + 1 < ntok) dst[k0 + 1] = v.y;
        if (k0 + 2 < ntok) dst[k0 + 2] = v.z;
        if (k0 + 3 < ntok) dst[k0 + 3] = v.w;
    }
}
// returns true when the entry now holds the ids (false: too many ids, marked not cacheable)
__device__ __forceinline__ bool memo_publish(const EncodeWorkspace &ws, uint32_t slot, const MemoKey &key, const uint32_t *buf,
                                             uint32_t ntok, uint32_t h6, uint32_t *status) {
    MemoEntry *e = ws.memo + slot;
    if (ntok > (uint32_t)kMemoTokens || h6 > 0xFFFFFFu) { st_release_u32(&e->meta, 0xFFFFFFFFu); return false; }
    e->tail_a = key.tail_a; e->tail_b = key.tail_b; e->tail_last = key.tail_last;
    for (uint32_t k = 0; k < ntok; ++k) e->tok[k] = buf[k];
    st_release_u32(&e->meta, (ntok + 1) | (h6 << 8));
    atomicAdd(&status[kStatusMemoTypes], 1u);
    return true;
}

// ---- warp-parallel decoupled look-back: all 32 lanes of one warp call this -----------------------------------------
// Publishes the tile's aggregate, then inspects 32 predecessors per step until a tile with a published inclusive
// prefix is found.  Returns (in every lane) the exclusive prefix of `tile`.
__device__ __forceinline__ uint64_t tile_prefix_warp(uint64_t *tile_state, uint32_t tile, uint64_t aggregate, uint32_t *err) {
    const uint32_t lane = threadIdx.x & 31;
    if (tile == 0) {
        if (lane == 0) st_relaxed_u64(&tile_state[0], (kTilePrefix << 62) | aggregate);
        return 0;
    }
    if (lane == 0) st_relaxed_u64(&tile_state[tile], (kTileAggregate << 62) | aggregate);
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
enum : uint32_t { kWordNone = 0u, kWordHit = 1u, kWordRecompute = 2u, kWordLong = 3u, kWordLongB = 4u };

// ---- the tile kernel ---------------------------------------------------------------------------------------------------
// Enc provides
//   uint32_t encode_short(const uint8_t *p, uint32_t nbytes, uint32_t *buf /*thread-local, kShortBytes*/, uint32_t &h6) const
//   static constexpr bool kCoopLong
//   kCoopLong == false:  uint32_t long_count(p, nbytes) const;  void long_emit(p, nbytes, uint32_t *dst, uint32_t cap, uint32_t &h6) const
//   kCoopLong == true :  uint32_t encode_long_coop(p, nbytes, bufA, bufB, uint32_t **result, uint32_t *sh_scan, uint32_t *sh_misc) const
template <class Enc>
__global__ void __launch_bounds__(kThreads)
encode_tiles_kernel(Enc enc, const uint8_t *__restrict__ arena, const uint32_t *__restrict__ word_off, uint32_t n_words,
                    uint32_t *__restrict__ out_ids, uint64_t out_cap, uint32_t *__restrict__ out_tok_off, uint32_t tok_base,
                    EncodeWorkspace ws, uint32_t *status) {
    __shared__ uint32_t compact[kCompactTokens];            // 32 KB: the tile's ids in output order
    __shared__ uint32_t s_off[kTileWords + 1];              // the tile's word offsets
    __shared__ uint32_t sh_scan[33];
    __shared__ uint32_t sh_misc[8];
    __shared__ uint64_t sh_base;
    const uint32_t tid = threadIdx.x;
    const uint32_t arena_end = word_off[n_words];
    const bool tok_off_vec = out_tok_off && (((uintptr_t)out_tok_off & 15) == 0);
    uint32_t h6 = 0;
    uint32_t next_ticket = 0;
    if (tid == 0) next_ticket = atomicAdd(ws.ticket, 1u);
    uint32_t buf[kShortBytes];                              // thread-local scratch for one directly encoded word

    for (;;) {
        if (tid == 0) sh_misc[4] = next_ticket;
        __syncthreads();
        const uint32_t tile = sh_misc[4];
        if (tile >= ws.n_tiles) break;
        const uint32_t w_tile = tile * kTileWords;
        const uint32_t tile_words = min((uint32_t)kTileWords, n_words - w_tile);
        for (uint32_t i = tid; i <= tile_words; i += kThreads) s_off[i] = word_off[w_tile + i];
        __syncthreads();

        // ---- phase A: counts
        uint32_t kind[kWordsPerThread], ntok[kWordsPerThread], slot[kWordsPerThread];
        uint32_t count = 0; bool has_long = false;
#pragma unroll
        for (int j = 0; j < kWordsPerThread; ++j) {
            const uint32_t i = tid * kWordsPerThread + j;
            kind[j] = kWordNone; ntok[j] = 0; slot[j] = 0;
            if (i >= tile_words) continue;
            const uint32_t b0 = s_off[i], nbytes = s_off[i + 1] - b0;
            if (nbytes > (uint32_t)kShortBytes) {
                kind[j] = kWordLong; has_long = true;
                if constexpr (!Enc::kCoopLong) ntok[j] = enc.long_count(arena + b0, nbytes);
            } else {
                int m = kMemoMiss; uint32_t meta = 0; MemoKey key;
                if (ws.memo_mask && nbytes >= 1) { memo_key(arena, b0, nbytes, arena_end, key); m = memo_probe(ws, key, slot[j], meta); }
                if (m == kMemoHit) { kind[j] = kWordHit; ntok[j] = (meta & 0xFFu) - 1; h6 += meta >> 8; }
                else {
                    uint32_t my_h6 = 0;
                    ntok[j] = enc.encode_short(arena + b0, nbytes, buf, my_h6);
                    h6 += my_h6;
                    kind[j] = (m == kMemoClaimed && memo_publish(ws, slot[j], key, buf, ntok[j], my_h6, status)) ? kWordHit : kWordRecompute;
                }
            }
            count += ntok[j];
        }
        if constexpr (Enc::kCoopLong) {
            // long words: the whole CTA works on one word at a time in global scratch (rare)
            if (__syncthreads_or(has_long)) {
                for (uint32_t i = 0; i < tile_words; ++i) {
                    const uint32_t b0 = s_off[i], nbytes = s_off[i + 1] - b0;
                    if (nbytes <= (uint32_t)kShortBytes) continue;                      // CTA-uniform
                    if (tid == 0) {
                        const unsigned long long so = atomicAdd(ws.long_cursor, 2ull * nbytes);
                        sh_misc[6] = (uint32_t)(so >> 1); sh_misc[7] = (so + 2ull * nbytes <= ws.long_scratch_elems) ? 1u : 0u;
                    }
                    __syncthreads();
                    const unsigned long long so = (unsigned long long)sh_misc[6] << 1;
                    const bool fits = sh_misc[7] != 0;
                    __syncthreads();
                    uint32_t *res = nullptr; uint32_t c = 0;
                    if (fits) c = enc.encode_long_coop(arena + b0, nbytes, ws.long_scratch + so, ws.long_scratch + so + nbytes, &res, sh_scan, sh_misc);
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

        // ---- scan + look-back
        uint32_t total, excl = block_exclusive_scan(count, sh_scan, &total);
        if (tid < 32) { const uint64_t b = tile_prefix_warp(ws.tile_state, tile, total, &status[kStatusCode]); if (tid == 0) sh_base = b; }
        __syncthreads();
        const uint64_t base = sh_base;
        const bool use_compact = total <= (uint32_t)kCompactTokens;
        const bool fits_out = base + total <= out_cap;
        if (!fits_out && tid == 0) atomicExch(&status[kStatusCode], (uint32_t)SWT_ERR_CAPACITY);

        // ---- phase B: ids to their tile-local position
        uint32_t run = excl;
        uint32_t offs[kWordsPerThread];
#pragma unroll
        for (int j = 0; j < kWordsPerThread; ++j) {
            const uint32_t i = tid * kWordsPerThread + j;
            offs[j] = tok_base + (uint32_t)(base + run);
            if (kind[j] == kWordNone || !fits_out) { run += ntok[j]; continue; }
            uint32_t *dst = use_compact ? compact + run : out_ids + base + run;
            if (kind[j] == kWordHit) memo_copy(ws, slot[j], ntok[j], dst);
            else if (kind[j] == kWordRecompute) {
                uint32_t dummy = 0;
                const uint32_t b0 = s_off[i];
                const uint32_t n = enc.encode_short(arena + b0, s_off[i + 1] - b0, buf, dummy);
                for (uint32_t k = 0; k < n; ++k) dst[k] = buf[k];
            } else {
                const uint32_t b0 = s_off[i], nbytes = s_off[i + 1] - b0;
                if constexpr (!Enc::kCoopLong) enc.long_emit(arena + b0, nbytes, dst, ntok[j], h6);
                else {
                    const uint32_t *src = ws.long_scratch + ((unsigned long long)slot[j] << 1) + (kind[j] == kWordLongB ? nbytes : 0u);
                    for (uint32_t k = 0; k < ntok[j]; ++k) dst[k] = src[k];
                }
            }
            run += ntok[j];
        }
        if (out_tok_off) {
            const uint32_t i0 = tid * kWordsPerThread;
            if (tok_off_vec && i0 + kWordsPerThread <= tile_words)
                *reinterpret_cast<uint4 *>(out_tok_off + w_tile + i0) = make_uint4(offs[0], offs[1], offs[2], offs[3]);
            else {
#pragma unroll
                for (int j = 0; j < kWordsPerThread; ++j) if (i0 + j < tile_words) out_tok_off[w_tile + i0 + j] = offs[j];
            }
        }
        __syncthreads();
        // take the next ticket now so that its latency hides behind the store of this tile.  (It must not be taken any
        // earlier: a tile that holds a ticket without running delays the look-back of every later tile.)
        if (tid == 0) next_ticket = atomicAdd(ws.ticket, 1u);
        if (use_compact && fits_out) for (uint32_t i = tid; i < total; i += kThreads) out_ids[base + i] = compact[i];
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
    if (!grid) grid = encode_grid((const void *)encode_tiles_kernel<Enc>, kThreads);
    const int g = (int)std::min<uint64_t>((uint64_t)grid, ws.n_tiles);
    encode_tiles_kernel<Enc><<<g, kThreads, 0, st>>>(enc, d_arena, d_word_off, n_words, d_out_ids, out_cap, d_out_tok_off, tok_base, ws, d_status);
    SWT_CUDA_OK(cudaGetLastError());
    return SWT_OK;
}
#endif  // __CUDACC__

}  // namespace swt
