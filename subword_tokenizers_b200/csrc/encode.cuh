// encode.cuh -- the tile skeleton shared by the two tokenize kernels (HP-1 FastBPE, HP-2 FastWP).
//
// Data flow of one launch (north-star subsystem 1: packed word-offset/byte arena):
//
//   arena bytes + u32 word offsets  --(one tile = kTileWords consecutive words per CTA)-->
//   per-word encode (tokens staged in shared memory, column per thread)                -->
//   CTA exclusive scan of the per-word token counts                                    -->
//   decoupled look-back over a 64-bit tile-state array (single pass, no second kernel) -->
//   compact token ids + u32 token offsets written at their final position.
//
// Tiles are handed out by an atomic ticket so the grid can be persistent (kNumSMs x resident CTAs)
// and so that the look-back never waits on a CTA that has not started.
#pragma once
#include "common.cuh"

namespace swt {

constexpr int kTileWords = 256;        // words per tile == threads per CTA
constexpr int kShortBytes = 32;        // words up to this many bytes take the shared-memory fast path

// status words written by the encode kernels
enum { kStatusCode = 0, kStatusTokens = 1, kStatusH6 = 2, kStatusTokensHi = 3, kStatusWords = 4 };

struct EncodeWorkspace {
    uint64_t *tile_state;   // n_tiles
    uint32_t *ticket;       // 1
    uint32_t *long_cursor;  // 1 (BPE: allocation cursor into long_scratch, in u32 units)
    uint32_t *long_scratch; // 2 x long bytes (BPE symbol ping-pong buffers for long words)
    uint64_t long_scratch_elems;
    uint32_t n_tiles;
};

size_t encode_workspace_layout(uint32_t n_words, uint64_t long_bytes, void *base, EncodeWorkspace *ws);

}  // namespace swt
