// bpe_train.cu -- HP-3: the merge loop of NaiveBPE.train (reference source/bpe.py:88-111) on sm_100a.
//
// HBM layout (all inside one caller-owned workspace):
//   sym[n_slots]        u32  the word table: symbols of all word types back to back, in first-occurrence
//                            order.  bit 31 marks the first symbol of a type; 0xFFFFFFFF is a dead slot
//                            (a merged word is compacted to the front of its own slot range, the tail dies).
//   word_of[n_slots]    u32  slot -> type index (read only where a merge happens)
//   start[n_types+1]    u32  type -> first slot;  freq[n_types] i64
//   table[cap]          16 B {key = left<<32|right, i64 count}: the pair-frequency map of bpe.py:90-95, kept
//                            INCREMENTALLY instead of being recounted every step
//   delta[2*vmax+2]     i64  per-step count deltas: L[x] (pairs (x,a)->(x,z)), R[y] ((b,y)->(z,y)), ZZ, M
//   symbol strings            length / offset / rolling hash per symbol + code arena, so that a merge whose
//                            concatenation already exists reuses that symbol (bpe.py:103)
//
// One step = select (argmax + first-occurrence tie-break) | merge (mark + in-place apply + deltas) |
// update (fold deltas into the table).  The only full-table passes are streaming reads: the argmax over the
// pair table, the mark scan over sym[], and -- only on a count tie -- an ascending, early-exit scan that
// finds the tied pair occurring first (Counter.most_common(1) returns the first-inserted maximum, bpe.py:102).
//
// Multi-GPU: types are sharded; L/R/ZZ/M are summed across ranks (NCCL all-reduce by the caller) and every
// rank folds the same global deltas into its replica of the table, so replicas stay identical.
#include <algorithm>
#include <cstdlib>
#include <map>
#include <vector>

#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace swt { extern int g_train_timing; }

namespace swt {

constexpr uint32_t kStart = 0x80000000u;
constexpr uint32_t kHole = 0xFFFFFFFFu;
constexpr uint64_t kEmptyKey = ~0ull;
constexpr uint64_t kNoPos = ~0ull;
constexpr uint32_t kMaxDenseAlpha = 4096;       // up to here the initial pair counts go through a dense n_alpha^2 array (one all-reduce
                                                // when sharded); larger alphabets (CJK, multilingual) count straight into the pair table
constexpr uint32_t kChunkShift = 12;            // mark scan granularity: 4096 slots per chunk
constexpr uint32_t kFiltBits = 16384;           // per-chunk pair filter: one bit per hashed pair that (may) occur in the chunk
constexpr uint32_t kBlkShift = 8;               // argmax cache, level 1: 256 table slots (4 KB) per block, refreshed by one warp
constexpr uint32_t kGrpShift = 8;               // level 2: 256 blocks (65,536 slots) per group

enum Halt : uint32_t { kRun = 0, kDoneVocab = 1, kDoneNoPairs = 2, kNeedGrow = 3, kRecordFull = 4,
                       kErrInternal = 16, kErrCharArena = 17, kErrSymbols = 18, kErrTableFull = 19, kErrScoreRange = 20, kErrPeerTimeout = 21 };

struct PairEntry { uint64_t key; long long count; };

struct TrainState {                 // device resident, mutable
    uint32_t halt, n_recorded;
    uint64_t n_merges_total;
    long long vocab_size;
    uint64_t n_symbols, n_entries, table_cap, n_live;
    // per-step scratch
    long long max_count;
    uint32_t n_tied, worklist_n, tie_ticket, step_stamp;
    uint64_t cand_key, best_pos;
    uint32_t cur_a, cur_b, cur_z, cur_valid;
    uint64_t char_used;
    uint32_t n_dirty, n_touch_l, n_touch_r, n_gdirty;   // lengths of the dirty-block / dirty-group lists and of the touched-symbol lists
    double best_score;                              // WordPiece mode: the maximal score of this step
    uint64_t n_tie_steps;                           // steps whose maximum was attained by several pairs (first-occurrence scan needed)
    uint32_t n_tie_keys, n_tie_listed;              // the tied pairs themselves when there are at most kTieKeys of them (else 0); listed steps
    uint64_t tie_keys[32];
    uint32_t xchg_epoch, pad_xchg;                  // peer exchange: number of cross-GPU barriers passed so far
    unsigned long long xchg_wait_cycles;            // ... and the SM cycles spent waiting in them (skew between the ranks + NVLink latency)
    unsigned long long xchg_cycles[3];              // diagnostic: SM cycles inside k_xchg_cand / k_xchg_delta / thread 0 of k_accum_peer
};
constexpr uint32_t kTieKeys = 32;

// ---- peer-memory exchange of the sharded trainer (NVLink / NVSwitch P2P stores, no NCCL call on the per-step path) -----------------
// Every rank owns one exchange buffer that all ranks map (symmetric memory).  Layout of a buffer:
//   [0, 256)      u32 flags[kMaxPeers]: flags[s] = number of barriers rank s has reached (written by rank s, read by the owner)
//   [256, 512)    u64 cand_in[kMaxPeers][2]: tie-break candidate (first position, pair) of every rank
//   [512, ...)    delta inboxes, two generations (barrier parity) x kMaxPeers source ranks x `stride` u64 words:
//                 header {n_l, n_r, zz, m}, then n_l + n_r entries {symbol, i64 delta}
// A rank PUSHES its data into the inbox of every rank (itself included), fences to system scope, raises its flag in every buffer
// and waits until all flags of its own buffer reached the barrier number.  One barrier per step (two on a tie step).
constexpr uint32_t kMaxPeers = 8;
struct PeerXchg {
    uint32_t enabled, pad;
    uint64_t stride;                                // u64 words of one source rank's delta inbox
    uint8_t *base[kMaxPeers];
};
__host__ __device__ __forceinline__ uint32_t *px_flags(uint8_t *b) { return reinterpret_cast<uint32_t *>(b); }
__host__ __device__ __forceinline__ uint64_t *px_cand(uint8_t *b) { return reinterpret_cast<uint64_t *>(b + 256); }
__host__ __device__ __forceinline__ uint64_t *px_delta(uint8_t *b, uint32_t parity, uint32_t src, uint64_t stride) {
    return reinterpret_cast<uint64_t *>(b + 512) + ((uint64_t)parity * kMaxPeers + src) * stride;
}
static inline uint64_t px_stride(uint32_t vmax) { return 4ull + 4ull * vmax; }       // header + 2 lists of up to vmax entries of 2 words
static inline size_t px_bytes(uint32_t vmax) { return 512 + 2ull * kMaxPeers * px_stride(vmax) * 8ull; }

struct ArgPart { long long count; uint64_t key; uint32_t n_tied; uint32_t pad; };
struct DirtyLists { uint32_t *dirty, *dirty_list, *gdirty, *gdirty_list; };

struct TrainDev {
    // sizes
    uint64_t n_types, n_slots, slot_base;
    uint32_t n_alpha, vmax, record_cap, world, rank, n_parts, mode, tie_chunk;
    long long max_vocab;
    uint64_t char_cap, str_ht_cap;
    // arrays
    uint32_t *sym, *word_of, *start, *word_mark, *worklist;
    uint32_t *sym2, *word_of2, *start2;             // second copies for the compaction of the word table (swt_bpe_train_maintain)
    uint32_t *filt; uint32_t n_chunks, mark_group;    // pair filter, bit-major: filt[(bit >> 5) * n_chunks + chunk]; chunks per CTA group of k_mark
    long long *freq;
    long long *sfreq;                 // WordPiece mode: frequency of every symbol (wordpiece.py:78-81), kept incrementally
    PairEntry *table;                 // current table (changes on grow)
    ArgPart *blk;                     // per 256-slot block: cached (max count, the smallest key attaining it, how many attain it)
    ArgPart *grp;                     // per group of 256 blocks: the same over the group's blocks
    uint32_t *dirty;                  // block touched since its cache entry was computed
    uint32_t *dirty_list;             // the blocks whose flag went 0 -> 1 (what the next select has to refresh)
    uint32_t *gdirty, *gdirty_list;   // the same per group
    uint32_t *touch_l, *touch_r;      // symbols x / y whose L[x] / R[y] became non-zero in this step
    long long *delta;                 // L[vmax] | R[vmax] | ZZ | M
    long long *dense;                 // n_alpha^2 initial counts
    uint64_t *cand;                   // 2
    uint64_t *cand_gather;            // world x 2
    ArgPart *parts;
    uint32_t *rec_left, *rec_right, *rec_new; long long *rec_count;
    uint32_t *sym_len; uint64_t *sym_off, *sym_hash, *sym_pow; uint32_t *chars; uint32_t *str_ht;
    TrainState *st;
    PeerXchg px;
    __host__ __device__ DirtyLists dl() const { return DirtyLists{dirty, dirty_list, gdirty, gdirty_list}; }
};

constexpr uint64_t kHashP = 0x9E3779B97F4A7C15ull | 1ull;

// ---- pair table --------------------------------------------------------------------------------------------
__device__ __forceinline__ long long table_get(const PairEntry *tab, uint64_t cap, uint64_t key) {
    uint64_t h = mix64(key) & (cap - 1);
    for (;;) {
        uint64_t k = tab[h].key;
        if (k == key) return tab[h].count;
        if (k == kEmptyKey) return 0;
        h = (h + 1) & (cap - 1);
    }
}
__device__ __forceinline__ void table_add(PairEntry *tab, uint64_t cap, uint64_t key, long long d, TrainState *st, const DirtyLists &dl) {
    uint64_t h = mix64(key) & (cap - 1);
    for (uint64_t probes = 0; probes <= cap; ++probes) {
        uint64_t k = *(volatile uint64_t *)&tab[h].key;
        if (k == kEmptyKey) {
            uint64_t old = atomicCAS((unsigned long long *)&tab[h].key, (unsigned long long)kEmptyKey, (unsigned long long)key);
            if (old == kEmptyKey) { atomicAdd((unsigned long long *)&st->n_entries, 1ull); k = key; }
            else k = old;
        }
        if (k == key) {
            atomicAdd((unsigned long long *)&tab[h].count, (unsigned long long)d);
            const uint32_t b = (uint32_t)(h >> kBlkShift);
            if (*(volatile uint32_t *)&dl.dirty[b] == 0u && atomicExch(&dl.dirty[b], 1u) == 0u) {
                dl.dirty_list[atomicAdd(&st->n_dirty, 1u)] = b;
                const uint32_t g = b >> kGrpShift;
                if (*(volatile uint32_t *)&dl.gdirty[g] == 0u && atomicExch(&dl.gdirty[g], 1u) == 0u) dl.gdirty_list[atomicAdd(&st->n_gdirty, 1u)] = g;
            }
            return;
        }
        h = (h + 1) & (cap - 1);
    }
    atomicExch(&st->halt, (uint32_t)kErrTableFull);
}

// ---- per-chunk pair filter: lets the mark scan skip the chunks that cannot contain the pair --------------------------------
// One bit per hashed pair and chunk of 4096 slots (16,384 bits per chunk, so a chunk's ~2,500 distinct pairs fill ~15 % of them).
// Bits are only ever set (a stale bit costs one wasted chunk scan).  The layout is bit-major, filt[(bit >> 5) * n_chunks + chunk],
// so that the query of one pair over all chunks is a contiguous read of 4 B per chunk (94 KB at 10 M word types) instead of the
// 385 MB word table.  Round 1 kept a per-chunk SYMBOL bitmap, which stops filtering once both symbols of a pair are common.
__device__ __forceinline__ uint32_t filt_bit(uint64_t key) { return (uint32_t)(mix64(key) >> 20) & (kFiltBits - 1u); }
__device__ __forceinline__ void filt_set(const TrainDev &d, uint64_t slot, uint32_t a, uint32_t b) {
    const uint32_t bit = filt_bit(((uint64_t)a << 32) | b);
    uint32_t *w = &d.filt[(uint64_t)(bit >> 5) * d.n_chunks + (slot >> kChunkShift)];
    const uint32_t m = 1u << (bit & 31u);
    if (!(*(volatile uint32_t *)w & m)) atomicOr(w, m);
}

// ---- init -----------------------------------------------------------------------------------------------------
__global__ void k_init_words(TrainDev d, const uint32_t *__restrict__ syms, const uint64_t *__restrict__ off) {
    // one thread per type: copies its symbols, flags the first, fills word_of / start
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < d.n_types; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t b = off[t], e = off[t + 1];
        d.start[t] = (uint32_t)b;
        for (uint64_t i = b; i < e; ++i) {
            d.sym[i] = syms[i] | (i == b ? kStart : 0u); d.word_of[i] = (uint32_t)t;
            if (i + 1 < e) filt_set(d, i, syms[i], syms[i + 1]);
        }
        if (t == d.n_types - 1) d.start[d.n_types] = (uint32_t)e;
    }
}
__global__ void k_init_symbols(TrainDev d, long long initial_vocab) {
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < d.n_alpha; c += gridDim.x * blockDim.x) {
        d.sym_len[c] = 1; d.sym_off[c] = c; d.chars[c] = c; d.sym_hash[c] = (uint64_t)c + 1; d.sym_pow[c] = kHashP;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        TrainState *st = d.st;
        st->halt = kRun; st->n_recorded = 0; st->n_merges_total = 0; st->vocab_size = initial_vocab;
        st->n_symbols = d.n_alpha; st->n_entries = 0; st->n_live = d.n_slots; st->step_stamp = 0;
        st->n_tie_steps = 0; st->n_tie_listed = 0;
        st->char_used = d.n_alpha; st->cur_valid = 0; st->worklist_n = 0; st->n_dirty = 0; st->n_gdirty = 0; st->n_touch_l = 0; st->n_touch_r = 0;
    }
}
__global__ void k_count_dense(TrainDev d) {
    const uint64_t n = d.n_slots;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i + 1 < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t s = d.sym[i], nx = d.sym[i + 1];
        if (nx & kStart) continue;                       // next slot starts another type (or is dead)
        const uint32_t a = s & ~kStart;
        atomicAdd((unsigned long long *)&d.dense[(uint64_t)a * d.n_alpha + nx], (unsigned long long)d.freq[d.word_of[i]]);
    }
}
// large alphabets: the initial counts go straight into the pair table (this rank's types only; see export / import below)
__global__ void k_count_sparse(TrainDev d, uint64_t cap) {
    const uint64_t n = d.n_slots;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i + 1 < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t s = d.sym[i], nx = d.sym[i + 1];
        if (nx & kStart) continue;
        table_add(d.table, cap, ((uint64_t)(s & ~kStart) << 32) | nx, d.freq[d.word_of[i]], d.st, d.dl());
    }
}
// sharded training with a large alphabet: every rank lists its local (pair, count) entries, the lists are all-gathered by
// the caller and the entries of the OTHER ranks are added, so that all replicas of the table hold the global counts
__global__ void k_export_pairs(TrainDev d, uint64_t cap, uint64_t *out, uint64_t out_cap, unsigned long long *n_out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
        const PairEntry e = d.table[i];
        if (e.key == kEmptyKey || e.count == 0) continue;
        const unsigned long long k = atomicAdd(n_out, 1ull);
        if (k < out_cap) { out[2 * k] = e.key; out[2 * k + 1] = (uint64_t)e.count; }
    }
}
__global__ void k_import_pairs(TrainDev d, uint64_t cap, const uint64_t *in, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        table_add(d.table, cap, in[2 * i], (long long)in[2 * i + 1], d.st, d.dl());
}
__global__ void k_build_table(TrainDev d, uint64_t cap) {
    const uint64_t n = (uint64_t)d.n_alpha * d.n_alpha;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const long long c = d.dense[i];
        if (c != 0) table_add(d.table, cap, ((i / d.n_alpha) << 32) | (i % d.n_alpha), c, d.st, d.dl());
    }
}
__global__ void k_fill_u64(uint64_t *p, uint64_t n, uint64_t v, uint64_t stride_words) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i * stride_words] = v;
}
__global__ void k_fill_u32(uint32_t *p, uint64_t n, uint32_t v) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_rehash(const PairEntry *old_tab, uint64_t old_cap, TrainDev d, uint64_t new_cap) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < old_cap; i += (uint64_t)gridDim.x * blockDim.x) {
        const PairEntry e = old_tab[i];
        if (e.key != kEmptyKey && e.count != 0) table_add(d.table, new_cap, e.key, e.count, d.st, d.dl());
    }
}

// ---- maintenance between step batches (swt_bpe_train_maintain) ------------------------------------------------------------------
// Merged words are compacted to the front of their own slot range, so the word table keeps its initial extent while three quarters
// of it die; the pair filter only ever gains bits.  Every few hundred steps the host asks for (a) a rebuild of the filter from the
// live pairs and, once less than 60 % of the slots are live, (b) a compaction of the whole table (words keep their order, so slot
// order is still first-occurrence order): the mark and tie scans then shrink with the corpus.
__global__ void k_word_len(TrainDev d, uint32_t *len) {
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < d.n_types; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t s = d.start[t], e = d.start[t + 1];
        uint32_t n = 0;
        while (s + n < e && d.sym[s + n] != kHole) ++n;
        len[t] = n;
    }
}
__global__ void k_compact_words(TrainDev d, const uint32_t *len, uint32_t n_new) {
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < d.n_types; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t s = d.start[t], ns = d.start2[t], n = len[t];
        for (uint32_t k = 0; k < n; ++k) { d.sym2[ns + k] = d.sym[s + k]; d.word_of2[ns + k] = (uint32_t)t; }
        if (t == d.n_types - 1) { d.start2[d.n_types] = n_new; for (uint32_t k = 0; k < 8; ++k) d.sym2[n_new + k] = kHole; }
    }
}
__global__ void k_build_filter(TrainDev d) {
    const uint64_t n = d.n_slots;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i + 1 < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t s = d.sym[i], nx = d.sym[i + 1];
        if (s == kHole || (nx & kStart)) continue;                       // dead slot / last symbol of its type (kHole has the start bit)
        filt_set(d, i, s & ~kStart, nx);
    }
}

// single rank: the delta vectors of the previous step are cleared through its touched-symbol lists (k_update reads every delta from
// two threads, so it cannot clear them itself); called by every thread of the select kernel before the lists are reset
__device__ __forceinline__ void clear_touched_deltas(const TrainDev &d, const TrainState *st) {
    if (d.world != 1 && !d.px.enabled) return;          // (NCCL exchange: k_update clears the dense vectors itself)
    long long *L = d.delta, *R = d.delta + d.vmax;
    const uint32_t nl = st->n_touch_l, nr = st->n_touch_r;
    for (uint32_t i = threadIdx.x; i < nl; i += blockDim.x) L[d.touch_l[i]] = 0;
    for (uint32_t i = threadIdx.x; i < nr; i += blockDim.x) R[d.touch_r[i]] = 0;
}

// ---- WordPiece mode (NaiveWP.train, reference source/wordpiece.py:29-103) -----------------------------------------------
// Same word table, pair table, mark / apply / update as BPE.  What differs: the initial symbols are strings ("a", "##a",
// wordpiece.py:53-57), the merged token is a + b[2:] (:95), and the pair chosen at every step maximises
// score = pair_freq / (freq[a] * freq[b]) (:84-92) with the first-inserted pair winning ties.  Python evaluates the score
// as int / int, i.e. the correctly rounded double of the exact quotient; freq[a]*freq[b] < 2^53 is required here so that
// the IEEE division of the two exactly representable operands gives the same double (the trainer halts otherwise).
__global__ void k_init_symbols_wp(TrainDev d, const uint32_t *__restrict__ init_cps, const uint64_t *__restrict__ init_off) {
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < d.n_alpha; c += gridDim.x * blockDim.x) {
        const uint64_t b = init_off[c], e = init_off[c + 1];
        uint64_t h = 0;
        for (uint64_t i = b; i < e; ++i) { d.chars[i] = init_cps[i]; h = h * kHashP + ((uint64_t)init_cps[i] + 1); }
        d.sym_len[c] = (uint32_t)(e - b); d.sym_off[c] = b; d.sym_hash[c] = h; d.sym_pow[c] = 0;
        uint64_t slot = mix64(h) & (d.str_ht_cap - 1);                 // initial symbols are vocabulary members too
        while (atomicCAS(&d.str_ht[slot], 0u, c + 1) != 0u) slot = (slot + 1) & (d.str_ht_cap - 1);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) d.st->char_used = init_off[d.n_alpha];
}
__global__ void k_count_sfreq(TrainDev d) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < d.n_slots; i += (uint64_t)gridDim.x * blockDim.x)
        atomicAdd((unsigned long long *)&d.sfreq[d.sym[i] & ~kStart], (unsigned long long)d.freq[d.word_of[i]]);
}
__device__ __forceinline__ double wp_score(const TrainDev &d, uint64_t key, long long count, TrainState *st) {
    const long long fa = d.sfreq[key >> 32], fb = d.sfreq[key & 0xFFFFFFFFull];
    const unsigned long long prod = (unsigned long long)fa * (unsigned long long)fb;
    if (fa <= 0 || fb <= 0 || __umul64hi((unsigned long long)fa, (unsigned long long)fb) != 0 || prod >= (1ull << 53) || count >= (1ll << 53)) {
        atomicExch(&st->halt, (uint32_t)kErrScoreRange);
        return 0.0;
    }
    return (double)count / (double)prod;                                // IEEE division == Python's correctly rounded int / int
}
__device__ __forceinline__ void score_combine(double &c, uint64_t &k, uint32_t &n, double c2, uint64_t k2, uint32_t n2) {
    if (c2 > c) { c = c2; k = k2; n = n2; }
    else if (c2 == c) { n += n2; if (k2 < k) k = k2; }
}
__global__ void __launch_bounds__(256) k_wp_argmax_partial(TrainDev d) {
    TrainState *st = d.st;
    if (st->halt) return;
    const uint64_t cap = st->table_cap;
    double c = 0.0; uint64_t k = kEmptyKey; uint32_t n = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
        const PairEntry e = d.table[i];
        if (e.key != kEmptyKey && e.count > 0) score_combine(c, k, n, wp_score(d, e.key, e.count, st), e.key, 1u);
    }
    for (int o = 16; o > 0; o >>= 1) {
        double c2 = __shfl_xor_sync(0xffffffffu, c, o); uint64_t k2 = __shfl_xor_sync(0xffffffffu, k, o);
        uint32_t n2 = __shfl_xor_sync(0xffffffffu, n, o);
        score_combine(c, k, n, c2, k2, n2);
    }
    __shared__ double sc[8]; __shared__ uint64_t sk[8]; __shared__ uint32_t sn[8];
    if ((threadIdx.x & 31) == 0) { sc[threadIdx.x >> 5] = c; sk[threadIdx.x >> 5] = k; sn[threadIdx.x >> 5] = n; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) score_combine(c, k, n, sc[w], sk[w], sn[w]);
        d.parts[blockIdx.x] = ArgPart{(long long)__double_as_longlong(c), k, n, 0};
    }
}
__global__ void __launch_bounds__(256) k_wp_select(TrainDev d) {
    TrainState *st = d.st;
    if (st->halt) return;
    clear_touched_deltas(d, st);
    double c = 0.0; uint64_t k = kEmptyKey; uint32_t n = 0;
    for (uint32_t i = threadIdx.x; i < d.n_parts; i += blockDim.x) { ArgPart p = d.parts[i]; score_combine(c, k, n, __longlong_as_double(p.count), p.key, p.n_tied); }
    for (int o = 16; o > 0; o >>= 1) {
        double c2 = __shfl_xor_sync(0xffffffffu, c, o); uint64_t k2 = __shfl_xor_sync(0xffffffffu, k, o);
        uint32_t n2 = __shfl_xor_sync(0xffffffffu, n, o);
        score_combine(c, k, n, c2, k2, n2);
    }
    __shared__ double sc[8]; __shared__ uint64_t sk[8]; __shared__ uint32_t sn[8];
    if ((threadIdx.x & 31) == 0) { sc[threadIdx.x >> 5] = c; sk[threadIdx.x >> 5] = k; sn[threadIdx.x >> 5] = n; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) score_combine(c, k, n, sc[w], sk[w], sn[w]);
        st->n_dirty = 0; st->n_gdirty = 0; st->n_touch_l = 0; st->n_touch_r = 0;
        st->best_score = c; st->max_count = n ? 1 : 0; st->n_tied = n; st->cand_key = k;
        st->worklist_n = 0; st->tie_ticket = 0; st->best_pos = kNoPos; st->cur_valid = 0;
        // loop conditions of wordpiece.py:68 and :74-75, then the capacity gates (checked before any mutation)
        if (st->vocab_size >= d.max_vocab) st->halt = kDoneVocab;
        else if (n == 0) st->halt = kDoneNoPairs;
        else if (st->n_recorded >= d.record_cap) st->halt = kRecordFull;
        else if ((st->n_entries + 2 * st->n_symbols + 2) * 10 > st->table_cap * 7) st->halt = kNeedGrow;
        else if (st->n_symbols + 1 > d.vmax) st->halt = kErrSymbols;
    }
}

// ---- select: argmax over the table (bpe.py:102) ---------------------------------------------------------------------
__device__ __forceinline__ void arg_combine(long long &c, uint64_t &k, uint32_t &n, long long c2, uint64_t k2, uint32_t n2) {
    if (c2 > c) { c = c2; k = k2; n = n2; }
    else if (c2 == c) { n += n2; if (k2 < k) k = k2; }
}
// Two-level cache of the table maximum.  Level 1: one entry per 256-slot block (4 KB of table), refreshed by ONE WARP per block
// that the last merge touched (a merge moves a few hundred pair counts, each in its own block: round 1 refreshed 4096-slot
// blocks, 64 KB each, 42 us per step at 10 M types).  Level 2: one entry per group of 256 blocks, refreshed from the level-1
// entries of the groups that hold a refreshed block.  k_select then reduces cap / 65,536 group entries.
__global__ void __launch_bounds__(256) k_argmax_blocks(TrainDev d) {
    TrainState *st = d.st;
    if (st->halt) return;
    const uint32_t n_dirty = st->n_dirty;
    const uint32_t lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t q = warp; q < n_dirty; q += n_warps) {
        const uint32_t b = d.dirty_list[q];
        long long c = 0; uint64_t k = kEmptyKey; uint32_t n = 0;
        const PairEntry *base = d.table + ((uint64_t)b << kBlkShift);
#pragma unroll
        for (uint32_t i = 0; i < (1u << kBlkShift) / 32; ++i) {
            const PairEntry e = base[i * 32 + lane];
            if (e.key != kEmptyKey && e.count > 0) arg_combine(c, k, n, e.count, e.key, 1u);
        }
        for (int o = 16; o > 0; o >>= 1) {
            long long c2 = __shfl_xor_sync(0xffffffffu, c, o); uint64_t k2 = __shfl_xor_sync(0xffffffffu, k, o);
            uint32_t n2 = __shfl_xor_sync(0xffffffffu, n, o);
            arg_combine(c, k, n, c2, k2, n2);
        }
        if (lane == 0) { d.blk[b] = ArgPart{c, k, n, 0}; d.dirty[b] = 0; }
    }
}
__global__ void __launch_bounds__(256) k_argmax_groups(TrainDev d) {
    TrainState *st = d.st;
    if (st->halt) return;
    const uint32_t n_gdirty = st->n_gdirty;
    const uint32_t n_blk = (uint32_t)(st->table_cap >> kBlkShift);
    const uint32_t lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t q = warp; q < n_gdirty; q += n_warps) {
        const uint32_t g = d.gdirty_list[q];
        long long c = 0; uint64_t k = kEmptyKey; uint32_t n = 0;
        for (uint32_t i = lane; i < (1u << kGrpShift); i += 32) {
            const uint32_t b = (g << kGrpShift) + i;
            if (b < n_blk) { const ArgPart p = d.blk[b]; arg_combine(c, k, n, p.count, p.key, p.n_tied); }
        }
        for (int o = 16; o > 0; o >>= 1) {
            long long c2 = __shfl_xor_sync(0xffffffffu, c, o); uint64_t k2 = __shfl_xor_sync(0xffffffffu, k, o);
            uint32_t n2 = __shfl_xor_sync(0xffffffffu, n, o);
            arg_combine(c, k, n, c2, k2, n2);
        }
        if (lane == 0) { d.grp[g] = ArgPart{c, k, n, 0}; d.gdirty[g] = 0; }
    }
}
__global__ void __launch_bounds__(1024) k_select(TrainDev d, int from_parts) {
    TrainState *st = d.st;
    if (st->halt) return;
    clear_touched_deltas(d, st);
    long long c = 0; uint64_t k = kEmptyKey; uint32_t n = 0;
    // reduce either the per-CTA partials of k_argmax_full or the group level of the cache (k_argmax_blocks / k_argmax_groups)
    const ArgPart *src = from_parts ? d.parts : d.grp;
    const uint32_t n_src = from_parts ? d.n_parts : (uint32_t)((st->table_cap + (1ull << (kBlkShift + kGrpShift)) - 1) >> (kBlkShift + kGrpShift));
    for (uint32_t i = threadIdx.x; i < n_src; i += blockDim.x) { ArgPart p = src[i]; arg_combine(c, k, n, p.count, p.key, p.n_tied); }
    for (int o = 16; o > 0; o >>= 1) {
        long long c2 = __shfl_xor_sync(0xffffffffu, c, o); uint64_t k2 = __shfl_xor_sync(0xffffffffu, k, o);
        uint32_t n2 = __shfl_xor_sync(0xffffffffu, n, o);
        arg_combine(c, k, n, c2, k2, n2);
    }
    __shared__ long long sc[32]; __shared__ uint64_t sk[32]; __shared__ uint32_t sn[32];
    if ((threadIdx.x & 31) == 0) { sc[threadIdx.x >> 5] = c; sk[threadIdx.x >> 5] = k; sn[threadIdx.x >> 5] = n; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 32; ++w) arg_combine(c, k, n, sc[w], sk[w], sn[w]);
        st->n_dirty = 0; st->n_gdirty = 0;                  // every listed block / group was refreshed
        st->n_touch_l = 0; st->n_touch_r = 0;               // consumed by the previous step's k_update
        st->max_count = c; st->n_tied = n; st->cand_key = k;
        st->worklist_n = 0; st->tie_ticket = 0; st->best_pos = kNoPos; st->cur_valid = 0; st->n_tie_keys = 0;
        // loop conditions of bpe.py:88 and :98-99, then the capacity gates (checked before any mutation)
        if (st->vocab_size >= d.max_vocab) st->halt = kDoneVocab;
        else if (c <= 0) st->halt = kDoneNoPairs;
        else if (st->n_recorded >= d.record_cap) st->halt = kRecordFull;
        else if ((st->n_entries + 2 * st->n_symbols + 2) * 10 > st->table_cap * 7) st->halt = kNeedGrow;
        else if (st->n_symbols + 1 > d.vmax) st->halt = kErrSymbols;
        sc[0] = c; sn[0] = (st->halt == kRun && !from_parts && d.mode == 0 && n >= 2 && n <= kTieKeys) ? n : 0u;
    }
    __syncthreads();
    // A tie among a few pairs: list them (the blocks whose cached maximum equals the maximum hold them), so that the first-occurrence
    // scan can use the pair filter and compare keys instead of probing the table at every position.
    if (sn[0]) {
        // groups whose cached maximum is the maximum -> their blocks -> their slots: three short parallel phases
        const long long cmax = sc[0];
        const uint32_t n_blk = (uint32_t)(st->table_cap >> kBlkShift);
        __shared__ uint32_t s_grp[64], s_blk[64], s_ng, s_nb;
        if (threadIdx.x == 0) { s_ng = 0; s_nb = 0; }
        __syncthreads();
        for (uint32_t g = threadIdx.x; g < n_src; g += blockDim.x)
            if (d.grp[g].count == cmax) { const uint32_t q = atomicAdd(&s_ng, 1u); if (q < 64) s_grp[q] = g; }
        __syncthreads();
        const uint32_t ng = min(s_ng, 64u);
        for (uint32_t w = threadIdx.x; w < ng << kGrpShift; w += blockDim.x) {
            const uint32_t b = (s_grp[w >> kGrpShift] << kGrpShift) + (w & ((1u << kGrpShift) - 1u));
            if (b < n_blk && d.blk[b].count == cmax) { const uint32_t q = atomicAdd(&s_nb, 1u); if (q < 64) s_blk[q] = b; }
        }
        __syncthreads();
        const uint32_t nb = min(s_nb, 64u);
        for (uint32_t w = threadIdx.x; w < nb << kBlkShift; w += blockDim.x) {
            const PairEntry e = d.table[((uint64_t)s_blk[w >> kBlkShift] << kBlkShift) + (w & ((1u << kBlkShift) - 1u))];
            if (e.key != kEmptyKey && e.count == cmax) { const uint32_t q = atomicAdd(&st->n_tie_keys, 1u); if (q < kTieKeys) st->tie_keys[q] = e.key; }
        }
        // (more than 64 groups / blocks cannot happen with at most kTieKeys tied pairs; a short list simply fails the n_tie_keys == n_tied
        // test of the scans and the table-probing scan runs instead)
    }
}

// Small tables (<= 2^20 slots): a plain parallel pass over the whole table beats the block cache (few, large blocks).
__global__ void __launch_bounds__(256) k_argmax_full(TrainDev d) {
    TrainState *st = d.st;
    if (st->halt) return;
    const uint64_t cap = st->table_cap;
    long long c = 0; uint64_t k = kEmptyKey; uint32_t n = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
        const PairEntry e = d.table[i];
        if (e.key != kEmptyKey && e.count > 0) arg_combine(c, k, n, e.count, e.key, 1u);
    }
    for (int o = 16; o > 0; o >>= 1) {
        long long c2 = __shfl_xor_sync(0xffffffffu, c, o); uint64_t k2 = __shfl_xor_sync(0xffffffffu, k, o);
        uint32_t n2 = __shfl_xor_sync(0xffffffffu, n, o);
        arg_combine(c, k, n, c2, k2, n2);
    }
    __shared__ long long sc[8]; __shared__ uint64_t sk[8]; __shared__ uint32_t sn[8];
    if ((threadIdx.x & 31) == 0) { sc[threadIdx.x >> 5] = c; sk[threadIdx.x >> 5] = k; sn[threadIdx.x >> 5] = n; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) arg_combine(c, k, n, sc[w], sk[w], sn[w]);
        d.parts[blockIdx.x] = ArgPart{c, k, n, 0};
    }
}

// ---- select: tie-break scan.  Among the pairs whose count equals the maximum, the winner is the one whose first
// occurrence comes first in (type, position) order == ascending slot order.  Chunks are handed out in ascending
// order; a chunk that starts after an already found position is skipped, so the scan stops early.
// chunk of the tie scan (TrainDev::tie_chunk): 512 slots = 2 positions per thread, one L2 round trip deep.  Larger chunks make
// every thread issue its table probes one after the other: 10 M types, 31,955 merges: 512 -> 4.38 s, 2048 -> 4.56 s, 4096 -> 4.87 s,
// 16384 -> 7.50 s; train-5K: 512 -> 0.28 s, 4096 -> 0.69 s.
__global__ void __launch_bounds__(256) k_tie_scan(TrainDev d) {
    TrainState *st = d.st;
    if (st->halt || st->n_tied <= 1) return;
    const long long target = st->max_count;
    const double target_score = st->best_score;
    const uint64_t cap = st->table_cap;
    __shared__ uint32_t s_chunk, s_stop;
    __shared__ unsigned long long s_best;
    if (st->n_tie_keys == st->n_tied) return;                       // handled by k_tie_scan_listed
    for (;;) {
        if (threadIdx.x == 0) {
            s_chunk = atomicAdd(&st->tie_ticket, 1u); s_best = kNoPos;
            const uint64_t c = (uint64_t)s_chunk * d.tie_chunk;
            // stop when past the end, or when an earlier chunk already holds an occurrence
            s_stop = (c >= d.n_slots) || (*(volatile uint64_t *)&st->best_pos < d.slot_base + c);
        }
        __syncthreads();
        if (s_stop) break;
        const uint64_t c0 = (uint64_t)s_chunk * d.tie_chunk;
        uint64_t mine = kNoPos;
        for (uint64_t i = c0 + threadIdx.x; i < c0 + d.tie_chunk && i + 1 < d.n_slots; i += blockDim.x) {
            const uint32_t s = d.sym[i], nx = d.sym[i + 1];
            if (s == kHole || (nx & kStart)) continue;
            const uint64_t key = ((uint64_t)(s & ~kStart) << 32) | nx;
            const long long cnt = table_get(d.table, cap, key);
            const bool tied = d.mode == 0 ? cnt == target : (cnt > 0 && wp_score(d, key, cnt, st) == target_score);
            if (tied) { mine = i; break; }
        }
        if (mine != kNoPos) atomicMin(&s_best, (unsigned long long)mine);
        __syncthreads();
        if (threadIdx.x == 0 && s_best != kNoPos) atomicMin((unsigned long long *)&st->best_pos, (unsigned long long)(d.slot_base + s_best));
        __syncthreads();
    }
}
// The tied pairs are listed (k_select found at most kTieKeys of them): groups of 64 filter chunks in ascending order; a chunk is
// read only when the pair filter admits one of the pairs, and positions are compared with the listed keys (no table probes).
__global__ void __launch_bounds__(256) k_tie_scan_listed(TrainDev d) {
    TrainState *st = d.st;
    if (st->halt || st->n_tied <= 1 || st->n_tie_keys != st->n_tied) return;
    __shared__ uint32_t s_chunk, s_stop;
    __shared__ unsigned long long s_best;
    {
        // ---- the tied pairs are listed (k_select): groups of 64 filter chunks in ascending order; a chunk is read only when the
        // filter admits one of the pairs, and positions are compared with the listed keys (no table probes)
        __shared__ uint64_t s_keys[kTieKeys];
        __shared__ uint32_t s_row[kTieKeys], s_mask[kTieKeys], s_list[64], s_flag[64], s_n;
        const uint32_t nk = st->n_tie_keys;
        if (threadIdx.x < nk) {
            const uint64_t k = st->tie_keys[threadIdx.x];
            const uint32_t bit = filt_bit(k);
            s_keys[threadIdx.x] = k; s_row[threadIdx.x] = bit >> 5; s_mask[threadIdx.x] = 1u << (bit & 31u);
        }
        constexpr uint32_t kTieGroup = 8;                                                    // chunks per ticket: all CTAs busy at once
        const uint64_t group_slots = (uint64_t)kTieGroup << kChunkShift;
        const uint32_t lane = threadIdx.x & 31;
        for (;;) {
            __syncthreads();
            if (threadIdx.x == 0) {
                s_chunk = atomicAdd(&st->tie_ticket, 1u); s_best = kNoPos; s_n = 0;
                const uint64_t c = (uint64_t)s_chunk * group_slots;
                s_stop = (c >= d.n_slots) || (*(volatile uint64_t *)&st->best_pos < d.slot_base + c);
            }
            __syncthreads();
            if (s_stop) break;
            if (threadIdx.x < kTieGroup) {                                                   // one chunk per thread: any key's bit set?
                const uint32_t chunk = s_chunk * kTieGroup + threadIdx.x;
                bool f = false;
                if (chunk < d.n_chunks) for (uint32_t k = 0; k < nk && !f; ++k) f = (d.filt[(uint64_t)s_row[k] * d.n_chunks + chunk] & s_mask[k]) != 0;
                s_flag[threadIdx.x] = f ? 1u : 0u;
            }
            __syncthreads();
            if (threadIdx.x == 0) {                                                          // ascending list of the flagged chunks
                uint32_t n = 0;
                for (uint32_t q = 0; q < kTieGroup; ++q) if (s_flag[q]) s_list[n++] = s_chunk * kTieGroup + q;
                s_n = n;
            }
            __syncthreads();
            const uint32_t n_list = s_n;
            uint64_t mine = kNoPos;
            for (uint32_t q = 0; q < n_list; ++q) {
                const uint64_t c0 = (uint64_t)s_list[q] << kChunkShift;
#pragma unroll
                for (uint32_t it = 0; it < (1u << kChunkShift) / (256 * 4); ++it) {           // 4 slots per thread, 128-bit loads, all in flight
                    const uint64_t i = c0 + ((uint64_t)it * 256 + threadIdx.x) * 4;
                    uint4 v = make_uint4(kHole, kHole, kHole, kHole);
                    if (i < d.n_slots) v = *reinterpret_cast<const uint4 *>(d.sym + i);      // sym[] is padded with dead slots
                    uint32_t nxt = __shfl_down_sync(0xffffffffu, v.x, 1);
                    if (lane == 31) nxt = (i + 4 < d.n_slots) ? d.sym[i + 4] : kHole;
                    const uint32_t sy[5] = {v.x, v.y, v.z, v.w, nxt};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (sy[j] == kHole || (sy[j + 1] & kStart)) continue;
                        const uint64_t key = ((uint64_t)(sy[j] & ~kStart) << 32) | sy[j + 1];
                        for (uint32_t k = 0; k < nk; ++k) if (s_keys[k] == key) mine = min(mine, i + j);
                    }
                }
            }
            if (mine != kNoPos) atomicMin(&s_best, (unsigned long long)mine);
            __syncthreads();
            if (threadIdx.x == 0 && s_best != kNoPos) atomicMin((unsigned long long *)&st->best_pos, (unsigned long long)(d.slot_base + s_best));
        }
    }
}
__global__ void k_candidate(TrainDev d) {
    TrainState *st = d.st;
    uint64_t pos = kNoPos, key = kEmptyKey;
    if (!st->halt) {
        if (st->n_tied > 1) { st->n_tie_steps += 1; if (st->n_tie_keys == st->n_tied) st->n_tie_listed += 1; }
        if (st->n_tied <= 1) key = st->cand_key;
        else if (st->best_pos != kNoPos) {
            pos = st->best_pos;
            const uint64_t i = pos - d.slot_base;
            key = ((uint64_t)(d.sym[i] & ~kStart) << 32) | d.sym[i + 1];
        }
    }
    d.cand[0] = pos; d.cand[1] = key;
    d.cand_gather[2 * d.rank] = pos; d.cand_gather[2 * d.rank + 1] = key;
}

// ---- merge: name the new symbol (string identity, bpe.py:103), record the merge (bpe.py:104) ---------------------------
__global__ void __launch_bounds__(32) k_begin_merge(TrainDev d) {
    TrainState *st = d.st;
    if (st->halt) return;
    const uint32_t lane = threadIdx.x;
    uint64_t key = kEmptyKey;
    if (st->n_tied <= 1) key = st->cand_key;
    else {
        uint64_t best = kNoPos;
        for (uint32_t r = 0; r < d.world; ++r) if (d.cand_gather[2 * r] < best) { best = d.cand_gather[2 * r]; key = d.cand_gather[2 * r + 1]; }
        if (best == kNoPos) { if (lane == 0) st->halt = kErrInternal; return; }
    }
    const uint32_t a = (uint32_t)(key >> 32), b = (uint32_t)key;
    // BPE: merged = a + b (bpe.py:103).  WordPiece: merged = a + b[2:] (wordpiece.py:95).
    const uint32_t skip = d.mode == 1 ? min(2u, d.sym_len[b]) : 0u;
    const uint32_t la = d.sym_len[a], lb = d.sym_len[b] - skip, L = la + lb;
    const uint32_t *ca = d.chars + d.sym_off[a], *cb = d.chars + d.sym_off[b] + skip;
    uint64_t h;
    if (d.mode == 0) h = d.sym_hash[a] * d.sym_pow[b] + d.sym_hash[b];
    else { h = 0; for (uint32_t i = 0; i < L; ++i) h = h * kHashP + ((uint64_t)(i < la ? ca[i] : cb[i - la]) + 1); }
    uint64_t slot = mix64(h) & (d.str_ht_cap - 1);
    uint32_t z = kHole;
    for (;;) {
        const uint32_t id1 = d.str_ht[slot];
        if (id1 == 0) break;
        const uint32_t cnd = id1 - 1;
        if (d.sym_hash[cnd] == h && d.sym_len[cnd] == L) {
            const uint32_t *cc = d.chars + d.sym_off[cnd];
            bool diff = false;
            for (uint32_t i = lane; i < L; i += 32) { const uint32_t want = i < la ? ca[i] : cb[i - la]; if (cc[i] != want) diff = true; }
            if (!__any_sync(0xffffffffu, diff)) { z = cnd; break; }
        }
        slot = (slot + 1) & (d.str_ht_cap - 1);
    }
    if (z == kHole) {                                   // a string never seen before: the vocabulary grows
        if (st->char_used + L > d.char_cap) { if (lane == 0) st->halt = kErrCharArena; return; }
        z = (uint32_t)st->n_symbols;
        uint32_t *cz = d.chars + st->char_used;
        for (uint32_t i = lane; i < L; i += 32) cz[i] = i < la ? ca[i] : cb[i - la];
        __syncwarp();
        if (lane == 0) {
            d.sym_len[z] = L; d.sym_off[z] = st->char_used; d.sym_hash[z] = h; d.sym_pow[z] = d.sym_pow[a] * d.sym_pow[b];
            d.str_ht[slot] = z + 1;
            st->char_used += L; st->n_symbols += 1; st->vocab_size += 1;
        }
    }
    if (lane == 0) {
        const uint32_t r = st->n_recorded;
        d.rec_left[r] = a; d.rec_right[r] = b; d.rec_new[r] = z;
        d.rec_count[r] = d.mode == 0 ? st->max_count : table_get(d.table, st->table_cap, key);
        st->n_recorded = r + 1; st->n_merges_total += 1;
        st->cur_a = a; st->cur_b = b; st->cur_z = z; st->cur_valid = 1; st->step_stamp += 1;
    }
}

// ---- merge: mark the types that contain (a,b): the pair filter names the candidate chunks, which are then read with 128-bit loads ----
// A CTA owns groups of 64 consecutive chunks: 64 threads test the filter bit of one chunk each (one coalesced 256-byte read), the
// flagged chunks are compacted into shared memory and scanned by the whole CTA.
constexpr uint32_t kMarkGroup = 64;
__global__ void __launch_bounds__(256) k_mark(TrainDev d) {
    const TrainState *st = d.st;
    if (st->halt || !st->cur_valid) return;
    const uint32_t a = st->cur_a, b = st->cur_b, stamp = st->step_stamp;
    const uint64_t n = d.n_slots;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t bit = filt_bit(((uint64_t)a << 32) | b);
    const uint32_t *frow = d.filt + (uint64_t)(bit >> 5) * d.n_chunks;
    const uint32_t fmask = 1u << (bit & 31u);
    __shared__ uint32_t s_list[kMarkGroup], s_n;
    const uint32_t grp = d.mark_group, n_groups = (d.n_chunks + grp - 1) / grp;
    for (uint32_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        if (threadIdx.x < grp) {
            const uint32_t chunk = g * grp + threadIdx.x;
            if (chunk < d.n_chunks && (frow[chunk] & fmask)) s_list[atomicAdd(&s_n, 1u)] = chunk;
        }
        __syncthreads();
        const uint32_t n_list = s_n;
        for (uint32_t q = 0; q < n_list; ++q) {
            const uint64_t c0 = (uint64_t)s_list[q] << kChunkShift;
            for (uint32_t it = 0; it < (1u << kChunkShift) / (256 * 4); ++it) {
                const uint64_t i = c0 + ((uint64_t)it * 256 + threadIdx.x) * 4;           // 4 slots per thread, 128-bit load
                uint4 v = make_uint4(kHole, kHole, kHole, kHole);
                if (i < n) v = *reinterpret_cast<const uint4 *>(d.sym + i);              // sym[] is padded with dead slots
                uint32_t nxt = __shfl_down_sync(0xffffffffu, v.x, 1);
                if (lane == 31) nxt = (i + 4 < n) ? d.sym[i + 4] : kHole;
                const bool m0 = (v.x & ~kStart) == a && v.y == b, m1 = (v.y & ~kStart) == a && v.z == b;
                const bool m2 = (v.z & ~kStart) == a && v.w == b, m3 = (v.w & ~kStart) == a && nxt == b;
                if (m0 | m1 | m2 | m3) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const bool m = k == 0 ? m0 : k == 1 ? m1 : k == 2 ? m2 : m3;
                        if (!m) continue;
                        const uint32_t w = d.word_of[i + k];
                        if (atomicExch(&d.word_mark[w], stamp) != stamp) d.worklist[atomicAdd(&d.st->worklist_n, 1u)] = w;
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ---- merge: greedy left-to-right replacement inside each marked type (bpe.py:25-48), emitting count deltas -------------
__global__ void __launch_bounds__(128) k_apply(TrainDev d) {
    TrainState *st = d.st;
    if (st->halt || !st->cur_valid) return;
    const uint32_t a = st->cur_a, b = st->cur_b, z = st->cur_z, n_work = st->worklist_n;
    long long *L = d.delta, *R = d.delta + d.vmax;
    long long zz = 0, m = 0, removed = 0;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_work; k += gridDim.x * blockDim.x) {
        const uint32_t w = d.worklist[k];
        const uint32_t ws = d.start[w], we = d.start[w + 1];
        const long long f = d.freq[w];
        uint32_t r = ws, o = ws, prev = 0; bool have_prev = false, last_merge = false;
        while (r < we) {
            uint32_t s = d.sym[r];
            if (s == kHole) break;
            s &= ~kStart;
            const uint32_t nx = (r + 1 < we) ? d.sym[r + 1] : kHole;
            if (s == a && nx == b) {
                if (have_prev) {
                    if (last_merge) zz += f;
                    else if (atomicAdd((unsigned long long *)&L[prev], (unsigned long long)f) == 0ull) d.touch_l[atomicAdd(&st->n_touch_l, 1u)] = prev;
                }
                m += f;
                if (have_prev) filt_set(d, o - 1, prev, z);          // the pair now in front of the merged symbol lives at slot o - 1
                d.sym[o] = z; prev = z; last_merge = true; r += 2;
            } else {
                if (have_prev && last_merge && atomicAdd((unsigned long long *)&R[s], (unsigned long long)f) == 0ull) d.touch_r[atomicAdd(&st->n_touch_r, 1u)] = s;
                // a pair that is new (z in front) or that moved left with the compaction may now belong to another chunk
                if (have_prev && (last_merge || o != r)) filt_set(d, o - 1, prev, s);
                d.sym[o] = s; prev = s; last_merge = false; r += 1;
            }
            have_prev = true; ++o;
        }
        for (uint32_t q = o; q < r; ++q) d.sym[q] = kHole;
        d.sym[ws] |= kStart;
        removed += (long long)(r - o);
    }
    if (zz) atomicAdd((unsigned long long *)&d.delta[2 * (uint64_t)d.vmax], (unsigned long long)zz);
    if (m) atomicAdd((unsigned long long *)&d.delta[2 * (uint64_t)d.vmax + 1], (unsigned long long)m);
    if (removed) atomicAdd((unsigned long long *)&st->n_live, (unsigned long long)(-removed));
}

// ---- update: fold the (globally summed) deltas into the replicated table ----------------------------------------------------
__global__ void __launch_bounds__(256) k_update(TrainDev d) {
    TrainState *st = d.st;
    if (st->halt || !st->cur_valid) return;
    const uint64_t a = st->cur_a, b = st->cur_b, z = st->cur_z;
    const uint64_t cap = st->table_cap;
    long long *L = d.delta, *R = d.delta + d.vmax;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    if (d.world == 1 || d.px.enabled) {
        // single rank, or peer exchange (k_accum_peer rebuilt the lists from all ranks): only the symbols whose delta became non-zero
        // have to be visited
        // one table update per thread (the -delta and the +delta of an entry are independent chains of random accesses);
        // L / R are cleared by k_clear_deltas afterwards
        const uint32_t nl = st->n_touch_l, nr = st->n_touch_r;
        for (uint32_t i = gtid; i < 2 * (nl + nr); i += gsz) {
            const uint32_t e = i >> 1; const bool plus = i & 1u;
            if (e < nl) {
                const uint64_t x = d.touch_l[e]; const long long l = L[x];
                table_add(d.table, cap, (x << 32) | (plus ? z : a), plus ? l : -l, st, d.dl());
            } else {
                const uint64_t y = d.touch_r[e - nl]; const long long r = R[y];
                table_add(d.table, cap, ((plus ? z : b) << 32) | y, plus ? r : -r, st, d.dl());
            }
        }
    } else {
        // sharded: the deltas were summed over ranks, the lists are rank-local -> visit every symbol
        for (uint32_t x = gtid; x < d.vmax; x += gsz) {
            const long long l = L[x], r = R[x];
            if (l) { table_add(d.table, cap, ((uint64_t)x << 32) | a, -l, st, d.dl()); table_add(d.table, cap, ((uint64_t)x << 32) | z, l, st, d.dl()); L[x] = 0; }
            if (r) { table_add(d.table, cap, (b << 32) | x, -r, st, d.dl()); table_add(d.table, cap, (z << 32) | x, r, st, d.dl()); R[x] = 0; }
        }
    }
    if (gtid == 0) {
        const long long zz = d.delta[2 * (uint64_t)d.vmax], m = d.delta[2 * (uint64_t)d.vmax + 1];
        if (zz) { table_add(d.table, cap, (b << 32) | a, -zz, st, d.dl()); table_add(d.table, cap, (z << 32) | z, zz, st, d.dl()); }
        if (m) table_add(d.table, cap, (a << 32) | b, -m, st, d.dl());
        if (d.mode == 1) { d.sfreq[a] -= m; d.sfreq[b] -= m; d.sfreq[z] += m; }      // wordpiece.py:78-81, incrementally
        d.delta[2 * (uint64_t)d.vmax] = 0; d.delta[2 * (uint64_t)d.vmax + 1] = 0;
    }
}

// ---- peer exchange kernels ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys_u32(uint32_t *p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Cross-GPU barrier number `epoch`, by the lanes of ONE warp after the caller's pushes were fenced: lane p raises this rank's flag in
// rank p's buffer and waits for rank p's flag in the own buffer.  Bounded spin: a rank that never arrives is an error, not a hang.
__device__ __forceinline__ bool px_barrier(const TrainDev &d, uint32_t epoch) {
    const uint32_t lane = threadIdx.x & 31;
    bool ok = true;
    if (lane < d.world) {
        st_release_sys_u32(px_flags(d.px.base[lane]) + d.rank, epoch);
        const uint32_t *mine = px_flags(d.px.base[d.rank]) + lane;
        uint32_t spins = 0;
        const long long c0 = clock64();
        while ((int32_t)(ld_acquire_sys_u32(mine) - epoch) < 0) { if (++spins > (1u << 27)) { ok = false; break; } }
        if (lane == (d.rank + 1) % d.world) atomicAdd(&d.st->xchg_wait_cycles, (unsigned long long)(clock64() - c0));   // one sample per barrier
    }
    return __all_sync(0xffffffffu, ok);
}
// tie steps only: every rank learns every rank's candidate (first position of a tied pair in its shard); replaces the all_gather
__global__ void __launch_bounds__(32) k_xchg_cand(TrainDev d) {
    TrainState *st = d.st;
    if (st->halt || st->n_tied <= 1) return;                        // replicated state: all ranks take the same branch
    const uint32_t lane = threadIdx.x, epoch = st->xchg_epoch + 1;
    const long long c0 = clock64();
    // (the release store of px_barrier, by the same lane, orders these stores before the flag)
    if (lane < d.world) { uint64_t *dst = px_cand(d.px.base[lane]) + 2 * d.rank; dst[0] = d.cand[0]; dst[1] = d.cand[1]; }
    const bool ok = px_barrier(d, epoch);
    if (lane == 0) { st->xchg_epoch = epoch; if (!ok) st->halt = kErrPeerTimeout; st->xchg_cycles[0] += (unsigned long long)(clock64() - c0); }
}
// every step: this rank's touched (symbol, delta) lists + ZZ / M go to every rank; replaces the all_reduce over 2 * vmax + 2 values
__global__ void __launch_bounds__(1024) k_xchg_delta(TrainDev d) {
    TrainState *st = d.st;
    if (st->halt || !st->cur_valid) return;
    const uint32_t nl = st->n_touch_l, nr = st->n_touch_r, epoch = st->xchg_epoch + 1, parity = epoch & 1u;
    long long *L = d.delta, *R = d.delta + d.vmax;
    const long long c0 = clock64();
    for (uint32_t p = 0; p < d.world; ++p) {
        uint64_t *dst = px_delta(d.px.base[p], parity, d.rank, d.px.stride);
        for (uint32_t i = threadIdx.x; i < nl + nr; i += blockDim.x) {
            const uint32_t sym = i < nl ? d.touch_l[i] : d.touch_r[i - nl];
            dst[4 + 2 * (uint64_t)i] = sym;
            dst[5 + 2 * (uint64_t)i] = (uint64_t)(i < nl ? L[sym] : R[sym]);
        }
        if (threadIdx.x == 0) { dst[0] = nl; dst[1] = nr; dst[2] = (uint64_t)d.delta[2 * (uint64_t)d.vmax]; dst[3] = (uint64_t)d.delta[2 * (uint64_t)d.vmax + 1]; }
    }
    __syncthreads();                                                // every read of L / R happened
    for (uint32_t i = threadIdx.x; i < nl + nr; i += blockDim.x) { if (i < nl) L[d.touch_l[i]] = 0; else R[d.touch_r[i - nl]] = 0; }
    if (threadIdx.x == 0) { d.delta[2 * (uint64_t)d.vmax] = 0; d.delta[2 * (uint64_t)d.vmax + 1] = 0; }
    // the CTA barrier orders every thread's pushes before the release stores of the flags (cumulativity): no system-scope fence per
    // thread (1024 MEMBAR.SYS behind a kernel that has just written the word table were the most expensive part of this kernel)
    __syncthreads();
    if (threadIdx.x < 32) {
        const bool ok = px_barrier(d, epoch);
        if (threadIdx.x == 0) {
            st->xchg_epoch = epoch; if (!ok) st->halt = kErrPeerTimeout; st->xchg_cycles[1] += (unsigned long long)(clock64() - c0);
            st->n_touch_l = 0; st->n_touch_r = 0;                   // the lists are rebuilt from the inboxes of all ranks (k_accum_peer)
        }
    }
}
// barrier only (swt_bpe_train_exchange_probe: measures what one cross-GPU barrier costs on this box)
__global__ void __launch_bounds__(32) k_xchg_probe(TrainDev d) {
    TrainState *st = d.st;
    const uint32_t epoch = st->xchg_epoch + 1;
    __threadfence_system();
    const bool ok = px_barrier(d, epoch);
    if (threadIdx.x == 0) { st->xchg_epoch = epoch; if (!ok) st->halt = kErrPeerTimeout; }
}
// after the barrier: the lists of ALL ranks (own inbox) are summed into the dense delta vectors and the lists of touched symbols are
// rebuilt, so that k_update folds every symbol into the pair table ONCE (applying each rank's list separately cost world_size times
// the table updates: 8 GPUs were slower than the NCCL all-reduce of round 1).  Deltas are sums of positive frequencies, so "the old
// value was zero" identifies the first contribution to a symbol exactly as in k_apply.
__global__ void __launch_bounds__(256) k_accum_peer(TrainDev d) {
    TrainState *st = d.st;
    if (st->halt || !st->cur_valid) return;
    const uint32_t parity = st->xchg_epoch & 1u;                    // the generation the barrier of this step closed
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    long long *L = d.delta, *R = d.delta + d.vmax;
    long long zz = 0, m = 0;
    const long long c0 = clock64();
    for (uint32_t r = 0; r < d.world; ++r) {
        const uint64_t *src = px_delta(d.px.base[d.rank], parity, r, d.px.stride);
        const uint32_t nl = (uint32_t)src[0], nr = (uint32_t)src[1];
        zz += (long long)src[2]; m += (long long)src[3];
        for (uint32_t i = gtid; i < nl + nr; i += gsz) {
            const uint32_t sym = (uint32_t)src[4 + 2 * (uint64_t)i];
            const unsigned long long v = src[5 + 2 * (uint64_t)i];
            if (i < nl) { if (atomicAdd((unsigned long long *)&L[sym], v) == 0ull) d.touch_l[atomicAdd(&st->n_touch_l, 1u)] = sym; }
            else if (atomicAdd((unsigned long long *)&R[sym], v) == 0ull) d.touch_r[atomicAdd(&st->n_touch_r, 1u)] = sym;
        }
    }
    if (gtid == 0) {
        d.delta[2 * (uint64_t)d.vmax] = zz; d.delta[2 * (uint64_t)d.vmax + 1] = m;      // (zeroed by k_xchg_delta after the push)
        st->xchg_cycles[2] += (unsigned long long)(clock64() - c0);
    }
}

__global__ void k_clear_halt(TrainState *st, uint32_t which, uint32_t reset_records) {
    if (st->halt == which) st->halt = kRun;
    if (reset_records) st->n_recorded = 0;
}
__global__ void k_set_table_cap(TrainState *st, uint64_t cap) { st->table_cap = cap; st->n_entries = 0; st->n_dirty = 0; st->n_gdirty = 0; }

}  // namespace swt

using namespace swt;

struct swt_bpe_trainer {
    swt_bpe_train_config cfg;
    TrainDev dev;
    int device;
    uint64_t table_cap;
    int grid_scan;      // persistent grid for streaming passes
    const uint64_t *d_off0 = nullptr;   // the caller's type offsets (layout of swt_bpe_train_read_corpus)
    uint64_t n_slots0 = 0;              // slots at create (the table is compacted as it shrinks)
    uint32_t n_compactions = 0;
    cudaGraphExec_t step_graph = nullptr;   // kStepsPerGraph whole steps, captured once (re-captured after a table grow)
    // swt_tune("train_timing", 1): every kernel is launched eagerly between two events and synchronised; totals on stderr at destroy
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::map<std::string, std::pair<double, uint64_t>> acc;
};
// launches one kernel of the step; in timing mode its device time is accumulated under `name`
#define TRAIN_LAUNCH(name, ...)                                                                                        \
    do {                                                                                                               \
        if (t->timing) cudaEventRecord(t->ev0, st);                                                                    \
        __VA_ARGS__;                                                                                                   \
        if (t->timing) {                                                                                               \
            cudaEventRecord(t->ev1, st); cudaEventSynchronize(t->ev1);                                                 \
            float ms = 0; cudaEventElapsedTime(&ms, t->ev0, t->ev1);                                                   \
            auto &a = t->acc[name]; a.first += ms; a.second += 1;                                                      \
        }                                                                                                              \
    } while (0)
static constexpr uint32_t kStepsPerGraph = 32;

static uint64_t choose_table_cap(const swt_bpe_train_config *cfg) {
    if (cfg->table_cap) return std::max<uint64_t>(next_pow2(cfg->table_cap), 1ull << kBlkShift);
    // large alphabet: the initial count inserts directly, so the table must hold every distinct pair of the corpus from the
    // start (at most one per symbol slot; sharded callers pass table_cap sized from the GLOBAL slot count)
    if (cfg->n_alpha > kMaxDenseAlpha)
        return next_pow2(std::max<uint64_t>(1ull << 16, 2 * cfg->n_slots_local + 8ull * (uint64_t)std::max<int64_t>(cfg->max_vocab, 1)));
    uint64_t a2 = (uint64_t)cfg->n_alpha * cfg->n_alpha;
    uint64_t want = std::max<uint64_t>(1ull << 16, 4 * std::min<uint64_t>(a2, 1ull << 24));
    want = std::max<uint64_t>(want, 8ull * (uint64_t)std::max<int64_t>(cfg->max_vocab, 1));
    return next_pow2(want);
}
static uint32_t vmax_of(const swt_bpe_train_config *cfg) {
    return (uint32_t)std::max<int64_t>(cfg->max_vocab, (int64_t)cfg->n_alpha) + 2;
}
static uint64_t char_cap_of(const swt_bpe_train_config *cfg) {
    uint64_t want = (uint64_t)vmax_of(cfg) * std::max<uint32_t>(cfg->max_word_len, 1);
    return std::min<uint64_t>(std::max<uint64_t>(want, 1ull << 20), 1ull << 26) + 4ull * cfg->n_alpha;
}

static size_t table_region_bytes(uint64_t cap) {
    const uint64_t n_blk = cap >> kBlkShift, n_grp = (n_blk >> kGrpShift) + 1;
    return align_up(cap * sizeof(PairEntry), 256) + align_up(n_blk * sizeof(ArgPart), 256) + 2 * align_up(n_blk * sizeof(uint32_t), 256) +
           align_up(n_grp * sizeof(ArgPart), 256) + 2 * align_up(n_grp * sizeof(uint32_t), 256);
}
static void table_region_carve(void *base, uint64_t cap, TrainDev *d) {
    uint8_t *p = (uint8_t *)base;
    const uint64_t n_blk = cap >> kBlkShift, n_grp = (n_blk >> kGrpShift) + 1;
    d->table = (PairEntry *)p; p += align_up(cap * sizeof(PairEntry), 256);
    d->blk = (ArgPart *)p; p += align_up(n_blk * sizeof(ArgPart), 256);
    d->dirty = (uint32_t *)p; p += align_up(n_blk * sizeof(uint32_t), 256);
    d->dirty_list = (uint32_t *)p; p += align_up(n_blk * sizeof(uint32_t), 256);
    d->grp = (ArgPart *)p; p += align_up(n_grp * sizeof(ArgPart), 256);
    d->gdirty = (uint32_t *)p; p += align_up(n_grp * sizeof(uint32_t), 256);
    d->gdirty_list = (uint32_t *)p;
}
static size_t train_layout(const swt_bpe_train_config *cfg, void *base, TrainDev *d, uint64_t table_cap) {
    Carver cv(base);
    const uint32_t vmax = vmax_of(cfg);
    d->n_types = cfg->n_types_local; d->n_slots = cfg->n_slots_local; d->slot_base = cfg->slot_base;
    d->n_alpha = cfg->n_alpha; d->vmax = vmax; d->record_cap = cfg->record_cap; d->world = cfg->world_size; d->rank = cfg->rank;
    d->max_vocab = cfg->max_vocab; d->n_parts = kNumSMs * 4;
    d->tie_chunk = 512;                                       // measured best from 23 k to 10 M types (profiles/README.md)
    if (const char *e = getenv("SWT_TIE_CHUNK")) { const long v = atol(e); if (v >= 256 && v <= 65536) d->tie_chunk = (uint32_t)v; }   // experiments
    d->char_cap = char_cap_of(cfg); d->str_ht_cap = next_pow2(4ull * vmax);
    d->st = cv.take<TrainState>(1);
    d->sym = cv.take<uint32_t>(d->n_slots + 8);
    d->n_chunks = (uint32_t)((d->n_slots + (1ull << kChunkShift) - 1) >> kChunkShift);
    d->filt = cv.take<uint32_t>((uint64_t)d->n_chunks * (kFiltBits / 32) + 1);
    d->mark_group = std::min<uint32_t>(kMarkGroup, std::max<uint32_t>(1u, d->n_chunks / (2 * kNumSMs)));      // small corpora: one chunk per CTA
    d->word_of = cv.take<uint32_t>(d->n_slots + 1);
    d->start = cv.take<uint32_t>(d->n_types + 1);
    d->sym2 = cv.take<uint32_t>(d->n_slots + 8);
    d->word_of2 = cv.take<uint32_t>(d->n_slots + 1);
    d->start2 = cv.take<uint32_t>(d->n_types + 1);
    d->word_mark = cv.take<uint32_t>(d->n_types + 1);
    d->worklist = cv.take<uint32_t>(d->n_types + 1);
    d->freq = cv.take<long long>(d->n_types + 1);
    d->sfreq = cv.take<long long>(vmax);
    d->mode = cfg->mode;
    d->delta = cv.take<long long>(2ull * vmax + 2);
    d->touch_l = cv.take<uint32_t>(vmax); d->touch_r = cv.take<uint32_t>(vmax);
    d->dense = cv.take<long long>(cfg->n_alpha <= kMaxDenseAlpha ? (uint64_t)cfg->n_alpha * cfg->n_alpha + 1 : 1);
    d->cand = cv.take<uint64_t>(2);
    d->cand_gather = cv.take<uint64_t>(2ull * cfg->world_size);
    d->parts = cv.take<ArgPart>(d->n_parts);
    d->rec_left = cv.take<uint32_t>(cfg->record_cap); d->rec_right = cv.take<uint32_t>(cfg->record_cap);
    d->rec_new = cv.take<uint32_t>(cfg->record_cap); d->rec_count = cv.take<long long>(cfg->record_cap);
    d->sym_len = cv.take<uint32_t>(vmax); d->sym_off = cv.take<uint64_t>(vmax); d->sym_hash = cv.take<uint64_t>(vmax);
    d->sym_pow = cv.take<uint64_t>(vmax); d->chars = cv.take<uint32_t>(d->char_cap); d->str_ht = cv.take<uint32_t>(d->str_ht_cap);
    uint8_t *region = cv.take<uint8_t>(table_region_bytes(table_cap));
    table_region_carve(region, table_cap, d);
    return cv.used();
}

SWT_API size_t swt_bpe_train_workspace_bytes(const swt_bpe_train_config *cfg) {
    if (!cfg) return 0;
    TrainDev d;
    return train_layout(cfg, nullptr, &d, choose_table_cap(cfg));
}
SWT_API size_t swt_bpe_train_table_bytes(uint64_t cap) { return table_region_bytes(next_pow2(cap)) + 256; }

SWT_API int swt_bpe_train_create(const swt_bpe_train_config *cfg, const uint32_t *d_syms, const uint64_t *d_off,
                                 const int64_t *d_freq, const uint32_t *d_init_cps, const uint64_t *d_init_off,
                                 void *d_workspace, size_t workspace_bytes, void *stream, swt_bpe_trainer **out) {
    SWT_REQUIRE(cfg && out && d_workspace, "NULL argument");
    SWT_REQUIRE(cfg->n_slots_local < 0xFFFFFFF0ull, "n_slots_local must be < 2^32");
    SWT_REQUIRE(cfg->n_types_local < 0xFFFFFFF0ull, "n_types_local must be < 2^32");
    SWT_REQUIRE(cfg->n_alpha >= 1 && cfg->n_alpha < (1u << 30), "n_alpha must be in [1, 2^30)");
    SWT_REQUIRE(cfg->world_size >= 1 && cfg->rank < cfg->world_size, "bad rank/world_size");
    SWT_REQUIRE(cfg->record_cap >= 1, "record_cap must be >= 1");
    SWT_REQUIRE(cfg->max_vocab < (1ll << 30), "max_vocab must be < 2^30");
    SWT_REQUIRE(cfg->n_types_local == 0 || (d_syms && d_off && d_freq), "NULL corpus pointer");
    SWT_REQUIRE(cfg->mode <= 1, "mode must be SWT_TRAIN_BPE or SWT_TRAIN_WP");
    SWT_REQUIRE(cfg->mode == 0 || (d_init_cps && d_init_off), "WordPiece mode needs the initial symbol strings");
    SWT_REQUIRE(cfg->mode == 0 || cfg->world_size == 1, "WordPiece training is single-rank");
    cudaStream_t st = (cudaStream_t)stream;
    int device = 0;
    SWT_CUDA_OK(cudaGetDevice(&device));
    swt_bpe_trainer *t = new swt_bpe_trainer();
    t->cfg = *cfg; t->device = device; t->table_cap = choose_table_cap(cfg);
    t->d_off0 = d_off; t->n_slots0 = cfg->n_slots_local;
    t->timing = g_train_timing != 0;
    if (t->timing) { cudaEventCreate(&t->ev0); cudaEventCreate(&t->ev1); }
    size_t need = train_layout(cfg, d_workspace, &t->dev, t->table_cap);
    if (need > workspace_bytes) { delete t; set_error("train workspace too small"); return SWT_ERR_CAPACITY; }
    int sms = kNumSMs; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    t->grid_scan = sms * 8;
    TrainDev &d = t->dev;
    // zero everything up to the pair table, then mark the table empty
    cudaError_t e = cudaMemsetAsync(d_workspace, 0, (uint8_t *)d.table - (uint8_t *)d_workspace, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d.table, 0, table_region_bytes(t->table_cap), st);
    if (e != cudaSuccess) { delete t; set_error(std::string("memset: ") + cudaGetErrorString(e)); return SWT_ERR_CUDA; }
    k_fill_u64<<<t->grid_scan, 256, 0, st>>>((uint64_t *)d.table, t->table_cap, kEmptyKey, 2);
    k_init_symbols<<<32, 256, 0, st>>>(d, cfg->initial_vocab);
    if (cfg->mode == 1) k_init_symbols_wp<<<32, 256, 0, st>>>(d, d_init_cps, d_init_off);
    k_set_table_cap<<<1, 1, 0, st>>>(d.st, t->table_cap);
    k_fill_u32<<<1, 32, 0, st>>>(d.sym + d.n_slots, 8, kHole);
    if (cfg->n_types_local) {
        k_init_words<<<t->grid_scan, 256, 0, st>>>(d, d_syms, d_off);
        e = cudaMemcpyAsync(d.freq, d_freq, cfg->n_types_local * sizeof(long long), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) { delete t; set_error(std::string("freq copy: ") + cudaGetErrorString(e)); return SWT_ERR_CUDA; }
        if (cfg->mode == 1) k_count_sfreq<<<t->grid_scan, 256, 0, st>>>(d);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) { delete t; set_error(std::string("init launch: ") + cudaGetErrorString(e)); return SWT_ERR_CUDA; }
    *out = t;
    return SWT_OK;
}

SWT_API void swt_bpe_train_destroy(swt_bpe_trainer *t) {
    if (!t) return;
    if (t->step_graph) cudaGraphExecDestroy(t->step_graph);
    if (t->timing) {
        double tot = 0;
        for (auto &kv : t->acc) tot += kv.second.first;
        for (auto &kv : t->acc)
            fprintf(stderr, "[swt train timing] %-18s %9.3f ms total %8llu launches %8.2f us each %5.1f %%\n", kv.first.c_str(), kv.second.first,
                    (unsigned long long)kv.second.second, 1e3 * kv.second.first / std::max<uint64_t>(kv.second.second, 1), 100 * kv.second.first / tot);
        cudaEventDestroy(t->ev0); cudaEventDestroy(t->ev1);
    }
    delete t;
}

SWT_API int swt_bpe_train_buffers(const swt_bpe_trainer *t, void **init_counts_ptr, uint64_t *init_counts_elems,
                                  void **cand_ptr, void **cand_gather_ptr, void **delta_ptr, uint64_t *delta_elems) {
    SWT_REQUIRE(t != nullptr, "NULL trainer");
    if (init_counts_ptr) *init_counts_ptr = t->dev.dense;
    if (init_counts_elems) *init_counts_elems = t->cfg.n_alpha <= kMaxDenseAlpha ? (uint64_t)t->cfg.n_alpha * t->cfg.n_alpha : 0;
    if (cand_ptr) *cand_ptr = t->dev.cand;
    if (cand_gather_ptr) *cand_gather_ptr = t->dev.cand_gather;
    if (delta_ptr) *delta_ptr = t->dev.delta;
    if (delta_elems) *delta_elems = 2ull * t->dev.vmax + 2;
    return SWT_OK;
}

SWT_API int swt_bpe_train_count_local(swt_bpe_trainer *t, void *stream) {
    SWT_REQUIRE(t != nullptr, "NULL trainer");
    if (t->dev.n_slots) {
        if (t->cfg.n_alpha <= kMaxDenseAlpha) k_count_dense<<<t->grid_scan, 256, 0, (cudaStream_t)stream>>>(t->dev);
        else k_count_sparse<<<t->grid_scan, 256, 0, (cudaStream_t)stream>>>(t->dev, t->table_cap);
    }
    SWT_CUDA_OK(cudaGetLastError());
    return SWT_OK;
}
SWT_API int swt_bpe_train_build_table(swt_bpe_trainer *t, void *stream) {
    SWT_REQUIRE(t != nullptr, "NULL trainer");
    if (t->cfg.n_alpha <= kMaxDenseAlpha) k_build_table<<<t->grid_scan, 256, 0, (cudaStream_t)stream>>>(t->dev, t->table_cap);
    SWT_CUDA_OK(cudaGetLastError());
    return SWT_OK;
}
SWT_API int swt_bpe_train_export_pairs(swt_bpe_trainer *t, uint64_t *d_out, uint64_t out_cap_entries, uint64_t *d_n_out, void *stream) {
    SWT_REQUIRE(t && d_n_out && (d_out || out_cap_entries == 0), "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    SWT_CUDA_OK(cudaMemsetAsync(d_n_out, 0, sizeof(uint64_t), st));
    k_export_pairs<<<t->grid_scan, 256, 0, st>>>(t->dev, t->table_cap, d_out, out_cap_entries, (unsigned long long *)d_n_out);
    SWT_CUDA_OK(cudaGetLastError());
    return SWT_OK;
}
SWT_API int swt_bpe_train_import_pairs(swt_bpe_trainer *t, const uint64_t *d_in, uint64_t n_entries, void *stream) {
    SWT_REQUIRE(t && (d_in || n_entries == 0), "NULL argument");
    if (n_entries) k_import_pairs<<<t->grid_scan, 256, 0, (cudaStream_t)stream>>>(t->dev, t->table_cap, d_in, n_entries);
    SWT_CUDA_OK(cudaGetLastError());
    return SWT_OK;
}

SWT_API int swt_bpe_train_select(swt_bpe_trainer *t, void *stream) {
    SWT_REQUIRE(t != nullptr, "NULL trainer");
    cudaStream_t st = (cudaStream_t)stream;
    if (t->cfg.mode == 1) {                       // WordPiece: scores move with the symbol frequencies -> full pass
        TRAIN_LAUNCH("wp_argmax", k_wp_argmax_partial<<<t->dev.n_parts, 256, 0, st>>>(t->dev));
        TRAIN_LAUNCH("wp_select", k_wp_select<<<1, 256, 0, st>>>(t->dev));
    } else {
        if (t->table_cap <= (1ull << 20)) {       // small table: plain parallel pass
            TRAIN_LAUNCH("argmax_full", k_argmax_full<<<t->dev.n_parts, 256, 0, st>>>(t->dev));
            TRAIN_LAUNCH("select", k_select<<<1, 1024, 0, st>>>(t->dev, 1));
        } else {                                  // large table: refresh only the blocks (then groups) the last merge touched
            TRAIN_LAUNCH("argmax_blocks", k_argmax_blocks<<<kNumSMs, 256, 0, st>>>(t->dev));
            TRAIN_LAUNCH("argmax_groups", k_argmax_groups<<<32, 256, 0, st>>>(t->dev));
            TRAIN_LAUNCH("select", k_select<<<1, 1024, 0, st>>>(t->dev, 0));
        }
    }
    if (t->dev.n_slots) {
        TRAIN_LAUNCH("tie_scan_listed", k_tie_scan_listed<<<t->grid_scan / 2, 256, 0, st>>>(t->dev));
        TRAIN_LAUNCH("tie_scan", k_tie_scan<<<t->grid_scan / 2, 256, 0, st>>>(t->dev));
    }
    TRAIN_LAUNCH("candidate", k_candidate<<<1, 1, 0, st>>>(t->dev));
    SWT_CUDA_OK(cudaGetLastError());
    return SWT_OK;
}
SWT_API int swt_bpe_train_merge(swt_bpe_trainer *t, void *stream) {
    SWT_REQUIRE(t != nullptr, "NULL trainer");
    cudaStream_t st = (cudaStream_t)stream;
    TRAIN_LAUNCH("begin_merge", k_begin_merge<<<1, 32, 0, st>>>(t->dev));
    if (t->dev.n_slots) {
        TRAIN_LAUNCH("mark", k_mark<<<(int)std::min<uint32_t>((uint32_t)t->grid_scan, (t->dev.n_chunks + t->dev.mark_group - 1) / t->dev.mark_group), 256, 0, st>>>(t->dev));
        TRAIN_LAUNCH("apply", k_apply<<<t->grid_scan / 2, 128, 0, st>>>(t->dev));
    }
    SWT_CUDA_OK(cudaGetLastError());
    return SWT_OK;
}
SWT_API int swt_bpe_train_update(swt_bpe_trainer *t, void *stream) {
    SWT_REQUIRE(t != nullptr, "NULL trainer");
    cudaStream_t st = (cudaStream_t)stream;
    if (t->dev.px.enabled) TRAIN_LAUNCH("accum_peer", k_accum_peer<<<32, 256, 0, st>>>(t->dev));
    const int blocks = (t->cfg.world_size == 1 || t->dev.px.enabled) ? 32 : (int)std::min<uint64_t>((t->dev.vmax + 255) / 256, 4096);
    TRAIN_LAUNCH("update", k_update<<<blocks, 256, 0, st>>>(t->dev));
    SWT_CUDA_OK(cudaGetLastError());
    return SWT_OK;
}
// peer exchange (swt_bpe_train_set_peers): the two exchanges of a sharded step as kernels of the step itself
SWT_API int swt_bpe_train_exchange_candidates(swt_bpe_trainer *t, void *stream) {
    SWT_REQUIRE(t != nullptr && t->dev.px.enabled, "peer exchange is not set up");
    cudaStream_t st = (cudaStream_t)stream;
    TRAIN_LAUNCH("xchg_cand", k_xchg_cand<<<1, 32, 0, st>>>(t->dev));
    SWT_CUDA_OK(cudaGetLastError());
    return SWT_OK;
}
SWT_API int swt_bpe_train_exchange_deltas(swt_bpe_trainer *t, void *stream) {
    SWT_REQUIRE(t != nullptr && t->dev.px.enabled, "peer exchange is not set up");
    cudaStream_t st = (cudaStream_t)stream;
    TRAIN_LAUNCH("xchg_delta", k_xchg_delta<<<1, 1024, 0, st>>>(t->dev));
    SWT_CUDA_OK(cudaGetLastError());
    return SWT_OK;
}
SWT_API int swt_bpe_train_exchange_probe(swt_bpe_trainer *t, uint32_t n_rounds, void *stream) {
    SWT_REQUIRE(t != nullptr && t->dev.px.enabled, "peer exchange is not set up");
    for (uint32_t k = 0; k < n_rounds; ++k) k_xchg_probe<<<1, 32, 0, (cudaStream_t)stream>>>(t->dev);
    SWT_CUDA_OK(cudaGetLastError());
    return SWT_OK;
}
SWT_API size_t swt_bpe_train_peer_bytes(const swt_bpe_train_config *cfg) { return cfg ? px_bytes(vmax_of(cfg)) : 0; }
SWT_API int swt_bpe_train_set_peers(swt_bpe_trainer *t, void *const *peer_buffers, uint32_t n_peers) {
    SWT_REQUIRE(t && peer_buffers, "NULL argument");
    SWT_REQUIRE(n_peers == t->cfg.world_size && n_peers <= kMaxPeers, "one buffer per rank, at most 8 ranks");
    for (uint32_t p = 0; p < n_peers; ++p) { SWT_REQUIRE(peer_buffers[p] != nullptr, "NULL peer buffer"); t->dev.px.base[p] = (uint8_t *)peer_buffers[p]; }
    t->dev.px.stride = px_stride(t->dev.vmax);
    t->dev.px.enabled = 1;
    t->dev.cand_gather = px_cand(t->dev.px.base[t->cfg.rank]);       // k_begin_merge reads the candidates of all ranks from the own inbox
    if (t->step_graph) { cudaGraphExecDestroy(t->step_graph); t->step_graph = nullptr; }
    return SWT_OK;
}
static int enqueue_step(swt_bpe_trainer *t, void *stream) {
    int rc = swt_bpe_train_select(t, stream); if (rc) return rc;
    if (t->dev.px.enabled) { rc = swt_bpe_train_exchange_candidates(t, stream); if (rc) return rc; }
    rc = swt_bpe_train_merge(t, stream); if (rc) return rc;
    if (t->dev.px.enabled) { rc = swt_bpe_train_exchange_deltas(t, stream); if (rc) return rc; }
    return swt_bpe_train_update(t, stream);
}

SWT_API int swt_bpe_train_steps(swt_bpe_trainer *t, uint32_t n_steps, void *stream) {
    SWT_REQUIRE(t != nullptr, "NULL trainer");
    SWT_REQUIRE(t->cfg.world_size == 1 || t->dev.px.enabled, "swt_bpe_train_steps needs the peer exchange when sharded (swt_bpe_train_set_peers); "
                "otherwise use select/merge/update with collectives");
    cudaStream_t st = (cudaStream_t)stream;
    // The step is launch-bound on small corpora (9 short kernels), so kStepsPerGraph steps are captured into one CUDA
    // graph and replayed; every kernel is self-gating on the halt flag, so replaying past the end is harmless.
    static const bool no_graph = getenv("SWT_TRAIN_NO_GRAPH") != nullptr;        // experiments: eager launches
    if (!t->step_graph && n_steps >= kStepsPerGraph && st != nullptr && !t->timing && !no_graph) {
        cudaGraph_t g = nullptr;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            int rc = SWT_OK;
            for (uint32_t s = 0; s < kStepsPerGraph && rc == SWT_OK; ++s) rc = enqueue_step(t, stream);
            cudaError_t e = cudaStreamEndCapture(st, &g);
            if (rc == SWT_OK && e == cudaSuccess && g) {
                if (cudaGraphInstantiate(&t->step_graph, g, 0) != cudaSuccess) t->step_graph = nullptr;
            }
            if (g) cudaGraphDestroy(g);
            if (!t->step_graph) fprintf(stderr, "[swt] trainer: CUDA graph capture of the step failed (%s); falling back to eager launches\n", cudaGetErrorString(e));
            (void)cudaGetLastError();
        } else { fprintf(stderr, "[swt] trainer: stream capture unavailable; eager launches\n"); (void)cudaGetLastError(); }
    }
    uint32_t done = 0;
    if (t->step_graph) {
        for (; done + kStepsPerGraph <= n_steps; done += kStepsPerGraph) SWT_CUDA_OK(cudaGraphLaunch(t->step_graph, st));
    }
    for (; done < n_steps; ++done) { int rc = enqueue_step(t, stream); if (rc) return rc; }
    return SWT_OK;
}

SWT_API int swt_bpe_train_read(swt_bpe_trainer *t, uint32_t *h_left, uint32_t *h_right, uint32_t *h_new, int64_t *h_count,
                               swt_bpe_train_state *state, void *stream) {
    SWT_REQUIRE(t && state, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    TrainState hs;
    SWT_CUDA_OK(cudaMemcpyAsync(&hs, t->dev.st, sizeof(TrainState), cudaMemcpyDeviceToHost, st));
    SWT_CUDA_OK(cudaStreamSynchronize(st));
    const uint32_t n = hs.n_recorded;
    if (n) {
        SWT_REQUIRE(h_left && h_right && h_new, "record arrays are NULL");
        SWT_CUDA_OK(cudaMemcpyAsync(h_left, t->dev.rec_left, n * 4, cudaMemcpyDeviceToHost, st));
        SWT_CUDA_OK(cudaMemcpyAsync(h_right, t->dev.rec_right, n * 4, cudaMemcpyDeviceToHost, st));
        SWT_CUDA_OK(cudaMemcpyAsync(h_new, t->dev.rec_new, n * 4, cudaMemcpyDeviceToHost, st));
        if (h_count) SWT_CUDA_OK(cudaMemcpyAsync(h_count, t->dev.rec_count, n * 8, cudaMemcpyDeviceToHost, st));
    }
    k_clear_halt<<<1, 1, 0, st>>>(t->dev.st, kRecordFull, 1u);
    SWT_CUDA_OK(cudaStreamSynchronize(st));
    state->halt = hs.halt; state->n_recorded = n; state->n_merges_total = hs.n_merges_total; state->vocab_size = hs.vocab_size;
    state->n_symbols = hs.n_symbols; state->n_table_entries = hs.n_entries; state->table_cap = hs.table_cap; state->n_live_slots = hs.n_live;
    state->n_tie_steps = hs.n_tie_steps; state->n_tie_listed = hs.n_tie_listed;
    state->n_peer_barriers = hs.xchg_epoch; state->peer_wait_cycles = hs.xchg_wait_cycles;
    for (int k = 0; k < 3; ++k) state->peer_kernel_cycles[k] = hs.xchg_cycles[k];
    return SWT_OK;
}

SWT_API int swt_bpe_train_grow_table(swt_bpe_trainer *t, void *d_new_table, uint64_t new_cap, void *stream) {
    SWT_REQUIRE(t && d_new_table, "NULL argument");
    new_cap = next_pow2(new_cap);
    SWT_REQUIRE(new_cap > t->table_cap, "new table must be larger");
    cudaStream_t st = (cudaStream_t)stream;
    PairEntry *old_tab = t->dev.table; const uint64_t old_cap = t->table_cap;
    void *region = (void *)(((uintptr_t)d_new_table + 255) / 256 * 256);
    SWT_CUDA_OK(cudaMemsetAsync(region, 0, table_region_bytes(new_cap), st));
    table_region_carve(region, new_cap, &t->dev);
    t->table_cap = new_cap;
    if (t->step_graph) { cudaGraphExecDestroy(t->step_graph); t->step_graph = nullptr; }   // kernels captured the old table pointer
    k_fill_u64<<<t->grid_scan, 256, 0, st>>>((uint64_t *)t->dev.table, new_cap, kEmptyKey, 2);
    k_set_table_cap<<<1, 1, 0, st>>>(t->dev.st, new_cap);
    k_rehash<<<t->grid_scan, 256, 0, st>>>(old_tab, old_cap, t->dev, new_cap);
    k_clear_halt<<<1, 1, 0, st>>>(t->dev.st, kNeedGrow, 0u);
    SWT_CUDA_OK(cudaGetLastError());
    return SWT_OK;
}

SWT_API int swt_bpe_train_maintain(swt_bpe_trainer *t, uint64_t n_live_slots, int *kernels_changed, void *stream) {
    SWT_REQUIRE(t != nullptr, "NULL trainer");
    cudaStream_t st = (cudaStream_t)stream;
    TrainDev &d = t->dev;
    if (kernels_changed) *kernels_changed = 0;
    if (d.n_slots == 0) return SWT_OK;
    if (n_live_slots * 10 <= d.n_slots * 6 && d.n_slots >= (1u << 16)) {
        // ---- compaction: live length per type -> exclusive scan -> copy into the second buffers -> swap
        uint32_t *len = d.worklist;                                  // free between steps
        k_word_len<<<t->grid_scan, 256, 0, st>>>(d, len);
        size_t tmp_bytes = 0;
        SWT_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, len, d.start2, (int)d.n_types, st));
        void *tmp = nullptr;
        SWT_CUDA_OK(cudaMalloc(&tmp, tmp_bytes + 16));
        cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, len, d.start2, (int)d.n_types, st);
        uint32_t last[2] = {0, 0};
        if (e == cudaSuccess) e = cudaMemcpyAsync(&last[0], d.start2 + d.n_types - 1, 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&last[1], len + d.n_types - 1, 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        cudaFree(tmp);
        if (e != cudaSuccess) { set_error(std::string("compaction scan: ") + cudaGetErrorString(e)); return SWT_ERR_CUDA; }
        const uint32_t n_new = last[0] + last[1];
        k_compact_words<<<t->grid_scan, 256, 0, st>>>(d, len, n_new);
        std::swap(d.sym, d.sym2); std::swap(d.word_of, d.word_of2); std::swap(d.start, d.start2);
        d.n_slots = n_new;
        d.n_chunks = (uint32_t)((d.n_slots + (1ull << kChunkShift) - 1) >> kChunkShift);
        d.mark_group = std::min<uint32_t>(kMarkGroup, std::max<uint32_t>(1u, d.n_chunks / (2 * kNumSMs)));
        t->n_compactions += 1;
        if (t->step_graph) { cudaGraphExecDestroy(t->step_graph); t->step_graph = nullptr; }     // the captured kernels hold the old extents
        if (kernels_changed) *kernels_changed = 1;
    }
    // ---- pair filter rebuilt from the live pairs (stale bits of merged-away pairs cost chunk scans)
    SWT_CUDA_OK(cudaMemsetAsync(d.filt, 0, ((size_t)d.n_chunks * (kFiltBits / 32) + 1) * 4, st));
    k_build_filter<<<t->grid_scan, 256, 0, st>>>(d);
    SWT_CUDA_OK(cudaGetLastError());
    return SWT_OK;
}

SWT_API int swt_bpe_train_read_corpus(swt_bpe_trainer *t, uint32_t *h_syms, uint32_t *h_len, void *stream) {
    SWT_REQUIRE(t && h_syms && h_len, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t n = t->dev.n_slots, nt = t->dev.n_types;
    // h_syms keeps the CALLER's layout (type w at its original offset) although the device table may have been compacted
    std::vector<uint32_t> start(nt + 1), cur(n + 1);
    std::vector<uint64_t> off0(nt + 1);
    SWT_CUDA_OK(cudaMemcpyAsync(cur.data(), t->dev.sym, n * 4, cudaMemcpyDeviceToHost, st));
    SWT_CUDA_OK(cudaMemcpyAsync(start.data(), t->dev.start, (nt + 1) * 4, cudaMemcpyDeviceToHost, st));
    if (nt) SWT_CUDA_OK(cudaMemcpyAsync(off0.data(), t->d_off0, (nt + 1) * 8, cudaMemcpyDeviceToHost, st));
    SWT_CUDA_OK(cudaStreamSynchronize(st));
    for (uint64_t w = 0; w < nt; ++w) {
        uint32_t len = 0;
        const uint64_t o = off0[w] - off0[0];
        for (uint32_t i = start[w]; i < start[w + 1] && cur[i] != kHole; ++i) { h_syms[o + len] = cur[i] & ~kStart; ++len; }
        h_len[w] = len;
    }
    return SWT_OK;
}
