// types.cu -- word-type table for the trainers, built on the device from the pre-tokenized corpus.
//
// Replaces, in front of the merge loops, what the reference does on the host:
//     word_freqs = Counter(words)  /  dict in first-occurrence order       source/bpe.py:73-81, source/wordpiece.py:49-52
//     symbols of every word type (characters; "##"-prefixed after the first for WordPiece)   bpe.py:77-81, wordpiece.py:53-60
// Input is the packed word arena + offsets that swt_pretok_write produces (SWT_PRETOK_BERT mode); output is the trainer's
// input: word types in FIRST-OCCURRENCE order (the tie-break of both trainers depends on it), their frequencies, and their
// code points, plus presence bitmaps of the characters seen in first / later positions (the initial alphabet).
//
// Three phases, all data-parallel over the words:
//   1. insert: open-addressing table keyed by the word's bytes; a slot is claimed by storing the claimer's word index
//      (the arena is immutable, so "same type" = byte-compare with the slot's representative word, no publication
//      protocol needed); first occurrence = atomicMin of the word index, frequency = warp-aggregated atomicAdd.
//   2. number: word i is the first occurrence of its type iff slot.first == i; an exclusive scan of these flags numbers
//      the types in first-occurrence order.
//   3. write: per type its representative word, frequency, character count (scanned into symbol offsets), code points.
#include <algorithm>

#include "common.cuh"

namespace swt {
namespace {

constexpr uint32_t kEmpty32 = 0xFFFFFFFFu;
constexpr uint32_t kScanTile = 4096;                       // elements per scan tile (one CTA of 256 threads x 16)
enum { kTyCode = 0, kTyTypes = 1, kTySymsLo = 2, kTySymsHi = 3 };

struct TypeSlot { uint32_t rep, first; long long freq_m1; };   // memset 0xFF: rep = first = EMPTY, freq_m1 = -1 (freq - 1)
static_assert(sizeof(TypeSlot) == 16, "TypeSlot must be 16 bytes");

struct TypesWs {
    TypeSlot *slots; uint64_t n_slots;
    uint32_t *word_slot;          // n_words
    uint32_t *word_type;          // n_words: exclusive scan of the first-occurrence flags (= type id at first occurrences)
    unsigned long long *tile_a;   // scan scratch (n_words / kScanTile + 1)
    unsigned long long *sym_off;  // n_types_cap + 1 (filled by the second scan; copied out by swt_types_write)
    unsigned long long *tile_b;   // scan scratch for the second scan
    uint64_t n_types_cap;
};

size_t types_layout(uint64_t n_words, uint64_t max_types, void *base, TypesWs *ws) {
    Carver c(base);
    ws->n_slots = next_pow2(std::max<uint64_t>(max_types, 512) * 2);
    ws->n_types_cap = std::min<uint64_t>(max_types, n_words);
    ws->slots = c.take<TypeSlot>(ws->n_slots);
    ws->word_slot = c.take<uint32_t>(n_words + 1);
    ws->word_type = c.take<uint32_t>(n_words + 1);
    ws->tile_a = c.take<unsigned long long>(n_words / kScanTile + 2);
    ws->sym_off = c.take<unsigned long long>(ws->n_types_cap + 2);
    ws->tile_b = c.take<unsigned long long>(ws->n_types_cap / kScanTile + 2);
    return c.used();
}

__device__ __forceinline__ uint64_t hash_bytes(const uint8_t *p, uint32_t n) {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ n;
    uint32_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t v = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) v |= (uint64_t)p[i + k] << (8 * k);
        h = mix64(h ^ v);
    }
    uint64_t v = 0;
    for (uint32_t k = 0; i + k < n; ++k) v |= (uint64_t)p[i + k] << (8 * k);
    return mix64(h ^ v ^ 0xD6E8FEB86659FD93ull);
}
__device__ __forceinline__ bool same_bytes(const uint8_t *a, const uint8_t *b, uint32_t n) {
    for (uint32_t k = 0; k < n; ++k) if (a[k] != b[k]) return false;
    return true;
}

// ---- phase 1 -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) types_insert_kernel(const uint8_t *__restrict__ arena, const uint32_t *__restrict__ off, uint32_t n_words,
                                                           TypesWs ws, uint32_t *status) {
    const uint64_t mask = ws.n_slots - 1;
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t base = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31ull; base < n_words; base += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = base + lane;                        // the lanes of a warp hold consecutive words
        uint32_t slot = kEmpty32 - lane;                       // distinct sentinels for idle / failed lanes
        if (i < n_words) {
            const uint32_t b0 = off[i], nb = off[i + 1] - b0;
            uint64_t h = hash_bytes(arena + b0, nb) & mask;
            for (uint32_t probe = 0; probe < 4096; ++probe, h = (h + 1) & mask) {
                uint32_t rep = ws.slots[h].rep;
                if (rep == kEmpty32) { const uint32_t old = atomicCAS(&ws.slots[h].rep, kEmpty32, (uint32_t)i); rep = old == kEmpty32 ? (uint32_t)i : old; }
                if (rep == (uint32_t)i) { slot = (uint32_t)h; break; }
                const uint32_t r0 = off[rep];
                if (off[rep + 1] - r0 == nb && same_bytes(arena + r0, arena + b0, nb)) { slot = (uint32_t)h; break; }
            }
            if (slot > 0xFFFFFFDFu) atomicExch(&status[kTyCode], (uint32_t)SWT_ERR_CAPACITY);          // table too full
            ws.word_slot[i] = slot;
        }
        // one atomic per distinct slot and warp: the lowest lane of a group holds the lowest word index
        const uint32_t peers = __match_any_sync(0xffffffffu, slot);
        if (i < n_words && slot <= 0xFFFFFFDFu && lane == (uint32_t)(__ffs(peers) - 1)) {
            atomicMin(&ws.slots[slot].first, (uint32_t)i);
            atomicAdd((unsigned long long *)&ws.slots[slot].freq_m1, (unsigned long long)__popc(peers));
        }
    }
}

// ---- generic exclusive scan of 32-bit items into 64-bit prefixes: tile sums, scan of the sums (one CTA), apply --------------
template <class F>
__device__ __forceinline__ void scan_tile_sums(F item, uint64_t n, unsigned long long *tile_sum) {
    __shared__ unsigned long long sh[8];
    const uint64_t t0 = (uint64_t)blockIdx.x * kScanTile;
    unsigned long long s = 0;
    for (uint32_t k = threadIdx.x; k < kScanTile; k += blockDim.x) if (t0 + k < n) s += item(t0 + k);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned long long t = 0; for (int w = 0; w < 8; ++w) t += sh[w]; tile_sum[blockIdx.x] = t; }
}
__global__ void scan_sums_kernel(unsigned long long *tile_sum, uint64_t n_tiles, unsigned long long *total_out) {
    // one thread: the number of tiles is n / 4096
    unsigned long long run = 0;
    for (uint64_t t = 0; t < n_tiles; ++t) { const unsigned long long v = tile_sum[t]; tile_sum[t] = run; run += v; }
    *total_out = run;
}
template <class F, class G>
__device__ __forceinline__ void scan_apply(F item, G store, uint64_t n, const unsigned long long *tile_sum) {
    // each thread owns 16 consecutive items of the tile
    __shared__ unsigned long long sh[9];
    const uint64_t t0 = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * 16;
    uint32_t v[16]; unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) { v[k] = t0 + k < n ? item(t0 + k) : 0u; s += v[k]; }
    unsigned long long incl = s;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned long long u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += u; }
    if (lane == 31) sh[wid] = incl;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned long long r = 0; for (int w = 0; w < 8; ++w) { const unsigned long long x = sh[w]; sh[w] = r; r += x; } }
    __syncthreads();
    unsigned long long run = tile_sum[blockIdx.x] + sh[wid] + incl - s;
#pragma unroll
    for (int k = 0; k < 16; ++k) { if (t0 + k < n) store(t0 + k, run); run += v[k]; }
}

// ---- phase 2: number the types -----------------------------------------------------------------------------------------
struct FirstFlag {
    TypesWs ws;
    __device__ __forceinline__ uint32_t operator()(uint64_t i) const {
        const uint32_t s = ws.word_slot[i];
        return s <= 0xFFFFFFDFu && ws.slots[s].first == (uint32_t)i;
    }
};
__global__ void __launch_bounds__(256) types_flag_sums_kernel(TypesWs ws, uint64_t n_words) { scan_tile_sums(FirstFlag{ws}, n_words, ws.tile_a); }
__global__ void __launch_bounds__(256) types_flag_apply_kernel(TypesWs ws, uint64_t n_words) {
    uint32_t *out = ws.word_type;
    scan_apply(FirstFlag{ws}, [out](uint64_t i, unsigned long long p) { out[i] = (uint32_t)p; }, n_words, ws.tile_a);
}

// ---- phase 3: per-type records ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) types_collect_kernel(const uint8_t *__restrict__ arena, const uint32_t *__restrict__ off, uint32_t n_words,
                                                            TypesWs ws, uint32_t *type_word, long long *freq, uint32_t *n_chars, uint64_t type_cap) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t s = ws.word_slot[i];
        if (s > 0xFFFFFFDFu || ws.slots[s].first != (uint32_t)i) continue;
        const uint32_t t = ws.word_type[i];
        if (t >= type_cap) continue;
        type_word[t] = (uint32_t)i;
        freq[t] = ws.slots[s].freq_m1 + 1;
        uint32_t c = 0;
        for (uint32_t b = off[i]; b < off[i + 1]; ++b) c += (arena[b] & 0xC0u) != 0x80u;
        n_chars[t] = c;
    }
}
struct CharCount { const uint32_t *n; __device__ __forceinline__ uint32_t operator()(uint64_t t) const { return n[t]; } };
__global__ void __launch_bounds__(256) types_len_sums_kernel(const uint32_t *n_chars, uint64_t n_types, unsigned long long *tile) { scan_tile_sums(CharCount{n_chars}, n_types, tile); }
__global__ void __launch_bounds__(256) types_len_apply_kernel(const uint32_t *n_chars, uint64_t n_types, const unsigned long long *tile, unsigned long long *sym_off) {
    scan_apply(CharCount{n_chars}, [sym_off](uint64_t t, unsigned long long p) { sym_off[t] = p; }, n_types, tile);
}
__global__ void __launch_bounds__(256) types_symbols_kernel(const uint8_t *__restrict__ arena, const uint32_t *__restrict__ off, const uint32_t *type_word,
                                                            uint64_t n_types, const unsigned long long *sym_off, uint32_t *cps,
                                                            uint32_t *first_bitmap, uint32_t *later_bitmap) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_types; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t i = type_word[t];
        const uint32_t b0 = off[i], nb = off[i + 1] - b0;
        unsigned long long o = sym_off[t];
        for (uint32_t k = 0; k < nb;) {
            uint32_t adv; const uint32_t cp = utf8_decode(arena + b0 + k, nb - k, adv);
            cps[o++] = cp;
            if (cp < 0x110000u) atomicOr((k == 0 ? first_bitmap : later_bitmap) + (cp >> 5), 1u << (cp & 31));
            k += adv;
        }
    }
}
// code points -> symbol ids through a dense table (BPE: one id per character; WordPiece: first / later tables differ)
__global__ void __launch_bounds__(256) types_map_kernel(uint32_t *cps, uint64_t n_syms, const unsigned long long *sym_off, uint64_t n_types,
                                                        const uint32_t *lut_first, const uint32_t *lut_later, uint32_t n_lut) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_types; t += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long a = sym_off[t], b = t + 1 < n_types ? sym_off[t + 1] : n_syms;
        for (unsigned long long k = a; k < b; ++k) { const uint32_t cp = cps[k]; cps[k] = cp < n_lut ? (k == a ? lut_first : lut_later)[cp] : kEmpty32; }
    }
}

int grid_for(uint64_t n, int per_thread = 1) {
    return (int)std::min<uint64_t>((n / per_thread + 255) / 256 + 1, (uint64_t)kNumSMs * 16);
}

}  // namespace
}  // namespace swt

using namespace swt;

SWT_API size_t swt_types_workspace_bytes(uint64_t n_words, uint64_t max_types) {
    TypesWs ws;
    return types_layout(n_words, max_types, nullptr, &ws);
}

SWT_API int swt_types_count(const uint8_t *d_arena, const uint32_t *d_word_off, uint32_t n_words, uint64_t max_types, void *d_workspace,
                            size_t workspace_bytes, uint32_t *d_status, void *stream) {
    SWT_REQUIRE(d_word_off && d_workspace && d_status, "NULL argument");
    SWT_REQUIRE(n_words == 0 || d_arena, "d_arena is NULL");
    SWT_REQUIRE(n_words < 0xFFFFFF00u && max_types >= 1, "n_words must be < 2^32 - 256, max_types >= 1");
    cudaStream_t st = (cudaStream_t)stream;
    TypesWs ws;
    if (types_layout(n_words, max_types, d_workspace, &ws) > workspace_bytes) { set_error("types workspace too small"); return SWT_ERR_CAPACITY; }
    SWT_CUDA_OK(cudaMemsetAsync(d_status, 0, 8 * sizeof(uint32_t), st));
    if (n_words == 0) return SWT_OK;
    SWT_CUDA_OK(cudaMemsetAsync(ws.slots, 0xFF, ws.n_slots * sizeof(TypeSlot), st));
    types_insert_kernel<<<grid_for(n_words), 256, 0, st>>>(d_arena, d_word_off, n_words, ws, d_status);
    const uint64_t n_tiles = ((uint64_t)n_words + kScanTile - 1) / kScanTile;
    types_flag_sums_kernel<<<(unsigned)n_tiles, 256, 0, st>>>(ws, n_words);
    scan_sums_kernel<<<1, 1, 0, st>>>(ws.tile_a, n_tiles, ws.sym_off);                      // total -> sym_off[0] (scratch)
    types_flag_apply_kernel<<<(unsigned)n_tiles, 256, 0, st>>>(ws, n_words);
    SWT_CUDA_OK(cudaMemcpyAsync(d_status + kTyTypes, ws.sym_off, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    SWT_CUDA_OK(cudaGetLastError());
    return SWT_OK;
}

SWT_API int swt_types_write(const uint8_t *d_arena, const uint32_t *d_word_off, uint32_t n_words, uint64_t max_types, void *d_workspace,
                            size_t workspace_bytes, uint32_t n_types, uint32_t *d_type_word, int64_t *d_freq, uint32_t *d_n_chars,
                            uint64_t *d_sym_off, uint32_t *d_first_bitmap, uint32_t *d_later_bitmap, uint32_t *d_status, void *stream) {
    SWT_REQUIRE(d_word_off && d_workspace && d_status && d_type_word && d_freq && d_n_chars && d_sym_off && d_first_bitmap && d_later_bitmap,
                "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    TypesWs ws;
    if (types_layout(n_words, max_types, d_workspace, &ws) > workspace_bytes) { set_error("types workspace too small"); return SWT_ERR_CAPACITY; }
    SWT_REQUIRE(n_types <= ws.n_types_cap, "n_types above the capacity the workspace was sized for");
    SWT_CUDA_OK(cudaMemsetAsync(d_first_bitmap, 0, 0x110000 / 8, st));
    SWT_CUDA_OK(cudaMemsetAsync(d_later_bitmap, 0, 0x110000 / 8, st));
    SWT_CUDA_OK(cudaMemsetAsync(d_sym_off, 0, sizeof(uint64_t), st));
    if (n_types == 0) return SWT_OK;
    types_collect_kernel<<<grid_for(n_words), 256, 0, st>>>(d_arena, d_word_off, n_words, ws, d_type_word, (long long *)d_freq, d_n_chars, n_types);
    const uint64_t n_tiles = ((uint64_t)n_types + kScanTile - 1) / kScanTile;
    types_len_sums_kernel<<<(unsigned)n_tiles, 256, 0, st>>>(d_n_chars, n_types, ws.tile_b);
    scan_sums_kernel<<<1, 1, 0, st>>>(ws.tile_b, n_tiles, (unsigned long long *)d_sym_off + n_types);     // closing offset = total
    types_len_apply_kernel<<<(unsigned)n_tiles, 256, 0, st>>>(d_n_chars, n_types, ws.tile_b, (unsigned long long *)d_sym_off);
    SWT_CUDA_OK(cudaMemcpyAsync(d_status + kTySymsLo, d_sym_off + n_types, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    SWT_CUDA_OK(cudaGetLastError());
    return SWT_OK;
}

SWT_API int swt_types_symbols(const uint8_t *d_arena, const uint32_t *d_word_off, const uint32_t *d_type_word, uint32_t n_types,
                              const uint64_t *d_sym_off, uint32_t *d_cps_out, uint32_t *d_first_bitmap, uint32_t *d_later_bitmap, void *stream) {
    SWT_REQUIRE(d_word_off && d_type_word && d_sym_off && d_cps_out && d_first_bitmap && d_later_bitmap, "NULL argument");
    if (n_types == 0) return SWT_OK;
    types_symbols_kernel<<<grid_for(n_types), 256, 0, (cudaStream_t)stream>>>(d_arena, d_word_off, d_type_word, n_types,
                                                                               (const unsigned long long *)d_sym_off, d_cps_out, d_first_bitmap, d_later_bitmap);
    SWT_CUDA_OK(cudaGetLastError());
    return SWT_OK;
}

SWT_API int swt_types_map_symbols(uint32_t *d_cps_inout, uint64_t n_syms, const uint64_t *d_sym_off, uint32_t n_types, const uint32_t *d_lut_first,
                                  const uint32_t *d_lut_later, uint32_t n_lut, void *stream) {
    SWT_REQUIRE(d_cps_inout && d_sym_off && d_lut_first && d_lut_later, "NULL argument");
    if (n_types == 0) return SWT_OK;
    types_map_kernel<<<grid_for(n_types), 256, 0, (cudaStream_t)stream>>>(d_cps_inout, n_syms, (const unsigned long long *)d_sym_off, n_types,
                                                                           d_lut_first, d_lut_later, n_lut);
    SWT_CUDA_OK(cudaGetLastError());
    return SWT_OK;
}
