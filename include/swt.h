/*
 * swt.h -- C ABI of libswt.so, the B200 (sm_100a) implementation of the three hot paths of
 * phtryll/subword-tokenizers:
 *
 *   HP-1  FastBPE.encode_word / tokenize      reference source/bpe.py:205-249
 *   HP-2  FastWP.tokenize / matchloop         reference source/wordpiece.py:233-316, source/utils.py:66-139
 *   HP-3  NaiveBPE.train (used by FastBPE)    reference source/bpe.py:50-112, :25-48
 *
 * The reference is pure Python and has no FFI of its own; the boundary it offers
 * is the class surface NaiveBPE/FastBPE/NaiveWP/FastWP.  Each entry point below names the
 * reference function it replaces.  The Python classes in subword_tokenizers_b200/ bind these
 * with ctypes (INTEGRATION.md shows the stub a maintainer of the reference would add).
 *
 * Conventions
 *   - plain C types only; every function returns an int status (SWT_OK == 0) unless stated;
 *     swt_last_error() gives the message of the last failure on the calling thread.
 *   - "d_" pointers are device pointers owned by the caller (the Python host owns them as torch
 *     tensors); "h_" pointers are host pointers.  `stream` is a cudaStream_t passed as void*.
 *   - hot calls (encode, train_step*) allocate nothing and only enqueue work on `stream`;
 *     they are asynchronous unless stated.  *_create / *_host calls may allocate and synchronise.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Data layouts
 *   word arena     UTF-8 bytes of all words back to back (no separators)          uint8
 *   word offsets   n_words+1 byte offsets into the arena                          uint32  (arena < 4 GiB per call)
 *   token ids      compact, words in input order                                  uint32
 *   token offsets  n_words+1 offsets into the token ids (optional, may be NULL)   uint32
 *
 * Token id spaces
 *   BPE:  token = (symbol << 1) | continuation.  continuation=1 is the "##" prefix of
 *         symbols[1:] (bpe.py:240-241).  symbol is the caller's id for a string mentioned by the
 *         merge list, or SWT_BPE_UNKNOWN_CP | code point for a character no merge mentions.
 *         The empty word yields the single token SWT_BPE_EMPTY_TOKEN (bpe.py:207-208).
 *   WP:   token = index into the caller's (sorted) vocabulary; n_vocab is "['UNK']"
 *         (wordpiece.py:257) and n_vocab+1 is "[UNK]" (wordpiece.py:149, via the "##" corner :260-261).
 */
#ifndef SWT_H
#define SWT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWT_ABI_VERSION 1

#define SWT_OK 0
#define SWT_ERR_CUDA 1          /* a CUDA runtime call failed (no device, OOM, launch failure) */
#define SWT_ERR_ARG 2           /* invalid argument */
#define SWT_ERR_CAPACITY 3      /* an output or workspace capacity was too small */
#define SWT_ERR_INTERNAL 4      /* device-side consistency check failed */
#define SWT_ERR_RANGE 5         /* a token id does not fit the requested 16-bit output (swt_encode_host16) */

#define SWT_BPE_UNKNOWN_CP 0x40000000u
#define SWT_BPE_EMPTY_TOKEN 0xFFFFFFFEu

/* ---- library ------------------------------------------------------------------------------ */
int swt_abi_version(void);
const char *swt_last_error(void);
int swt_device_count(int *count);
/* Process-wide experiment knobs (diagnostics; defaults are the measured best).  "memo_max_log2" (10..23, default 22): cap of the
 * word-type memo of an encode call; "memo_off" (0/1): disable the memo, every word takes the direct path; "bulk_store" (0/1):
 * cp.async.bulk copy-out in the emit pass; "warp_words": FastBPE warp-per-word threshold; "timing" (0/1): per-kernel times on
 * stderr.  A change of the memo knobs changes swt_encode_workspace_bytes: size workspaces after setting them. */
int swt_tune(const char *name, int value);

/* ---- HP-1: FastBPE rank table + encode ------------------------------------------------------ */
/*
 * swt_bpe_table_create  replaces  FastBPE._bpe_ranks = {pair: i ...}   (bpe.py:200, :257)
 *   merge k is (left[k], right[k]) -> merged[k]; ids are the caller's, canonical per distinct
 *   string; a pair listed twice keeps its LAST rank (dict semantics).  char_cp (ascending) /
 *   char_id map a code point to the id of its one-character symbol.
 *   The table (open-addressing hash, key = left<<32|right) is built on the host and uploaded to
 *   `device`.
 */
typedef struct swt_bpe_table swt_bpe_table;
int swt_bpe_table_create(const uint32_t *h_left, const uint32_t *h_right, const uint32_t *h_merged, uint32_t n_merges,
                         const uint32_t *h_char_cp, const uint32_t *h_char_id, uint32_t n_chars, int device,
                         swt_bpe_table **out);
void swt_bpe_table_destroy(swt_bpe_table *t);

/* Words longer than SWT_SHORT_WORD_BYTES leave the shared-memory fast path; `long_word_bytes` is the
 * total byte length of those words (pass the arena size when unknown).
 * swt_encode_workspace_bytes: bytes of device workspace an encode call needs (same for BPE and WP). */
#define SWT_SHORT_WORD_BYTES 32
size_t swt_encode_workspace_bytes(uint32_t n_words, uint64_t long_word_bytes);

/*
 * swt_bpe_encode  replaces  [tok for w in words for tok in FastBPE.encode_word(w)]   (bpe.py:205-249)
 *   d_status (8 x u32): [0] SWT_OK / SWT_ERR_CAPACITY / SWT_ERR_INTERNAL, [1] total tokens (low 32 bits,
 *   [3] high bits; also written to d_out_tok_off[n_words] when that is not NULL), [2] H6 events (WP);
 *   diagnostics: [4] word types published in the call's memo, [5] words that took the slow path.
 *   Asynchronous on `stream`; returns only argument/launch errors.
 */
int swt_bpe_encode(const swt_bpe_table *t, const uint8_t *d_arena, const uint32_t *d_word_off, uint32_t n_words,
                   uint64_t long_word_bytes, uint32_t *d_out_ids, uint64_t out_cap, uint32_t *d_out_tok_off,
                   void *d_workspace, size_t workspace_bytes, uint32_t *d_status, void *stream);

/* ---- HP-2: WordPiece trie + encode ------------------------------------------------------------ */
/*
 * swt_wp_trie_create  replaces  WPTrie_E2E(vocab)   (utils.py:75-139; insert :87-105, precompute :108-139)
 *   vocab token v is the code points h_vocab_cps[h_vocab_off[v] .. h_vocab_off[v+1]).
 *   h_alnum_bitmap: 0x110000 bits, bit cp = Python's chr(cp).isalnum() (utils.py:137,
 *   wordpiece.py:287-288 use the Python predicates; the bitmap is copied to the device).
 *   The trie, failure links and failure pops are computed on the host and flattened into
 *   device tables.  n_sharp_special/sharp_special give NaiveWP.encode_word("##")
 *   (wordpiece.py:260-261), which the caller evaluates once.
 */
typedef struct swt_wp_trie swt_wp_trie;
int swt_wp_trie_create(const uint32_t *h_vocab_cps, const uint64_t *h_vocab_off, uint32_t n_vocab,
                       const uint8_t *h_alnum_bitmap, const uint32_t *h_sharp_special, uint32_t n_sharp_special,
                       int device, swt_wp_trie **out);
void swt_wp_trie_destroy(swt_wp_trie *t);
/* nodes / edges / pops / nodes whose failure link is root_p (SURVEY.md §8 a7 figures) */
int swt_wp_trie_stats(const swt_wp_trie *t, uint64_t *n_nodes, uint64_t *n_edges, uint64_t *n_pops, uint64_t *n_rootp);

/*
 * swt_wp_encode  replaces  FastWP.tokenize   (wordpiece.py:233-270 + matchloop :291-316)
 *   Each input "word" is a whitespace-free chunk of the lower-cased text (the host splits on
 *   Python str.isspace); the kernel appends the virtual " " of wordpiece.py:248 itself.
 *   H6 extension: on a punctuation character that is not a child of the trie root
 *   the reference never terminates; this implementation emits "['UNK']" and advances one
 *   character, and counts such events in d_status[2].
 *   Status / outputs as swt_bpe_encode.
 */
int swt_wp_encode(const swt_wp_trie *t, const uint8_t *d_arena, const uint32_t *d_word_off, uint32_t n_words,
                  uint64_t long_word_bytes, uint32_t *d_out_ids, uint64_t out_cap, uint32_t *d_out_tok_off,
                  void *d_workspace, size_t workspace_bytes, uint32_t *d_status, void *stream);

/* ---- the Naive encoders ---------------------------------------------- */
/*
 * swt_bpe_encode_naive  replaces  NaiveBPE.encode_word  (bpe.py:114-132): the merge list replayed in order.  Same
 *   table, arguments and outputs as swt_bpe_encode.  Exact for merge lists in which no (left, right) pair is listed
 *   twice (the table keeps one rank per pair); the Python class keeps the host replay for lists with repeats.
 * swt_wp_encode_naive   replaces  NaiveWP.encode_word   (wordpiece.py:131-158): greedy longest prefix, "##" + rest,
 *   whole word -> "[UNK]" (id n_vocab + 1) when a piece has no match.  Same trie, arguments and outputs as swt_wp_encode;
 *   the input words are the BERT pre-tokenized words (wordpiece.py:176-178), not whitespace chunks.
 */
int swt_bpe_encode_naive(const swt_bpe_table *t, const uint8_t *d_arena, const uint32_t *d_word_off, uint32_t n_words,
                         uint64_t long_word_bytes, uint32_t *d_out_ids, uint64_t out_cap, uint32_t *d_out_tok_off,
                         void *d_workspace, size_t workspace_bytes, uint32_t *d_status, void *stream);
int swt_wp_encode_naive(const swt_wp_trie *t, const uint8_t *d_arena, const uint32_t *d_word_off, uint32_t n_words,
                        uint64_t long_word_bytes, uint32_t *d_out_ids, uint64_t out_cap, uint32_t *d_out_tok_off,
                        void *d_workspace, size_t workspace_bytes, uint32_t *d_status, void *stream);

/* ---- device-side pre-tokenization for FastWP -------------------------- */
/*
 * swt_pretok_*  replaces  text.lower().split()  i.e. `s = text.lower() + " "` (wordpiece.py:248) and the
 * whitespace skipping of FastWP.tokenize (wordpiece.py:266-269), on raw UTF-8 text resident on the device.
 * Output: the packed word arena (lower-cased bytes of the whitespace-free chunks) + n_words+1 u32 offsets,
 * exactly what swt_wp_encode consumes.
 *   lower_map[cp] for cp < n_lower: the lower-case code point; 0x80000000|i = one-to-many mapping stored at
 *   multi[i] = count, multi[i+1..] = code points (U+0130); 0x40000000 = U+03A3, resolved with the final-sigma
 *   rule of CPython (needs the Cased / Case_Ignorable bitmaps, 0x110000/8 bytes each, bit cp&7 of byte cp>>3;
 *   both NULL = the caller guarantees the text holds no U+03A3, else the call reports SWT_ERR_ARG).
 *   Whitespace is Python's str.isspace() set (29 code points, fixed in the kernel).
 * The text must be valid UTF-8 (lone surrogates in their 3-byte form are passed through), 16-byte aligned, in a
 * buffer readable up to the next multiple of 4 bytes, and < 4 GiB per call.
 * Two calls: swt_pretok_count fills d_status (8 x u32: [0] status, [1] n_words, [2]/[3] arena bytes lo/hi) so the
 * caller can size the outputs; swt_pretok_write (same text, same workspace, untouched in between) writes them.
 * mode SWT_PRETOK_BERT  replaces  pre_tokenize_str(example.lower())  of SubwordTokenizer.preprocessing (utils.py:26-29,
 * the Rust BertPreTokenizer): whitespace is Rust's char::is_whitespace (no U+001C-001F), and every punctuation
 * character (is_bert_punc: ASCII punctuation, or bit 29 = 0x20000000 of its lower_map entry) is a word of its own.
 */
#define SWT_PRETOK_PYTHON_SPLIT 0
#define SWT_PRETOK_BERT 1
typedef struct swt_pretok swt_pretok;
int swt_pretok_create(const uint32_t *lower_map, uint32_t n_lower, const uint32_t *multi, uint32_t n_multi,
                      const uint8_t *cased_bitmap, const uint8_t *ignorable_bitmap, int mode, int device, swt_pretok **out);
void swt_pretok_destroy(swt_pretok *p);
size_t swt_pretok_workspace_bytes(uint64_t n_text_bytes);
int swt_pretok_count(const swt_pretok *p, const uint8_t *d_text, uint64_t n_bytes, void *d_workspace, size_t workspace_bytes,
                     uint32_t *d_status, void *stream);
int swt_pretok_write(const swt_pretok *p, const uint8_t *d_text, uint64_t n_bytes, void *d_workspace, size_t workspace_bytes,
                     uint8_t *d_arena_out, uint64_t arena_cap, uint32_t *d_word_off_out, uint32_t *d_word_src_out, uint64_t word_cap,
                     uint32_t n_words, uint64_t n_out_bytes, uint32_t *d_status, void *stream);
/* d_word_src_out (n_words u32, may be NULL): byte position in the text at which each word starts, so that a caller who
 * concatenated many texts can cut the token stream per text (tokenize_batch; the harness row of SURVEY.md section 8f). */

/* ---- word-type table for the trainers, built on the device ------------------------------------------- */
/*
 * swt_types_*  replaces  `word_freqs = Counter(words)` in first-occurrence order and the per-type symbol lists in front
 * of the two merge loops (bpe.py:73-81, wordpiece.py:49-60), on the packed words that swt_pretok_write produced
 * (SWT_PRETOK_BERT).  Types come out in FIRST-OCCURRENCE order (both trainers break ties by it).
 *   swt_types_count   : hash-dedupes the words; d_status (8 x u32): [0] status (SWT_ERR_CAPACITY: more than max_types
 *                       distinct words -- call again with a larger max_types), [1] n_types.
 *   swt_types_write   : (same arena / workspace, untouched) per type: index of its first-occurrence word, frequency,
 *                       character count, exclusive symbol offsets (n_types + 1); d_status [2]/[3] = total symbols lo/hi.
 *                       Clears the two bitmaps.
 *   swt_types_symbols : the code points of every type at its symbol offset, and the presence bitmaps (0x110000 bits
 *                       each) of the characters seen in first / in later positions (the initial alphabet).
 *   swt_types_map_symbols : code points -> symbol ids in place through caller-built dense tables (BPE: the same table
 *                       twice; WordPiece: first-position and "##" tables), 0xFFFFFFFF for cp >= n_lut.
 */
size_t swt_types_workspace_bytes(uint64_t n_words, uint64_t max_types);
int swt_types_count(const uint8_t *d_arena, const uint32_t *d_word_off, uint32_t n_words, uint64_t max_types, void *d_workspace,
                    size_t workspace_bytes, uint32_t *d_status, void *stream);
int swt_types_write(const uint8_t *d_arena, const uint32_t *d_word_off, uint32_t n_words, uint64_t max_types, void *d_workspace,
                    size_t workspace_bytes, uint32_t n_types, uint32_t *d_type_word, int64_t *d_freq, uint32_t *d_n_chars,
                    uint64_t *d_sym_off, uint32_t *d_first_bitmap, uint32_t *d_later_bitmap, uint32_t *d_status, void *stream);
int swt_types_symbols(const uint8_t *d_arena, const uint32_t *d_word_off, const uint32_t *d_type_word, uint32_t n_types,
                      const uint64_t *d_sym_off, uint32_t *d_cps_out, uint32_t *d_first_bitmap, uint32_t *d_later_bitmap, void *stream);
int swt_types_map_symbols(uint32_t *d_cps_inout, uint64_t n_syms, const uint64_t *d_sym_off, uint32_t n_types, const uint32_t *d_lut_first,
                          const uint32_t *d_lut_later, uint32_t n_lut, void *stream);

/* ---- host-buffer entry points (what a non-Python integrator binds) ------------------------------- */
/*
 * Same results as the device-pointer calls above, but inputs and outputs are HOST buffers:
 * the call stages the corpus through the GPU in batches (H2D, encode, D2H overlapped on
 * separate streams) and returns when h_out_ids / h_out_tok_off are complete.
 * which: 0 = BPE (table is a swt_bpe_table*), 1 = WP (table is a swt_wp_trie*).
 * h_out_tok_off may be NULL.  n_tokens receives the total; h6_events may be NULL.
 */
typedef struct swt_pipeline swt_pipeline;
int swt_pipeline_create(int device, uint64_t batch_bytes, swt_pipeline **out);
void swt_pipeline_destroy(swt_pipeline *p);
/* h_word_off: n_words+1 u32 byte offsets (arena < 4 GiB per call; split larger corpora);
 * h_out_tok_off: n_words+1 u32 or NULL.  Pinned host buffers (swt_host_alloc) let the copies overlap. */
int swt_encode_host(swt_pipeline *p, int which, const void *table, const uint8_t *h_arena, const uint32_t *h_word_off,
                    uint64_t n_words, uint32_t *h_out_ids, uint64_t out_cap, uint32_t *h_out_tok_off,
                    uint64_t *n_tokens, uint64_t *h6_events);
/* The same call with 16-bit token ids in the host buffer: halves the device-to-host traffic, which bounds the
 * end-to-end rate (PCIe).  Usable when every token id is < 65536 (WordPiece: vocabulary + 2 <= 65536; BPE: symbol
 * ids < 32768 and no character outside the merge alphabet); otherwise the call fails with SWT_ERR_RANGE and the
 * caller falls back to swt_encode_host.  The reference returns a flat token list (wordpiece.py:270, bpe.py:249),
 * so h_out_tok_off == NULL is the like-for-like output. */
int swt_encode_host16(swt_pipeline *p, int which, const void *table, const uint8_t *h_arena, const uint32_t *h_word_off,
                      uint64_t n_words, uint16_t *h_out_ids16, uint64_t out_cap, uint32_t *h_out_tok_off,
                      uint64_t *n_tokens, uint64_t *h6_events);
/* FastWP.tokenize / FastBPE.tokenize from RAW TEXT in a host buffer (wordpiece.py:233-270 / bpe.py:245-249 end to end):
 * per batch H2D of the text, swt_pretok_count/_write, the encode call, D2H of the flat token ids (16-bit when
 * ids_16bit != 0, see swt_encode_host16).  which / table as in swt_encode_host; pretok must have the matching mode
 * (SWT_PRETOK_PYTHON_SPLIT for WP, SWT_PRETOK_BERT for BPE).  Batches are cut after ASCII whitespace.
 * n_words_out / h6_events may be NULL. */
int swt_tokenize_text_host(swt_pipeline *p, const swt_pretok *pretok, int which, const void *table, const uint8_t *h_text,
                           uint64_t n_bytes, void *h_out_ids, int ids_16bit, uint64_t out_cap, uint64_t *n_tokens,
                           uint64_t *n_words_out, uint64_t *h6_events);
/* Small calls: ONE short text per call -- the pattern of the reference's CLI, which calls tokenize() once per line
 * (cli.py:253-264; wordpiece.py:233-270 / bpe.py:245-249 end to end, and the Naive encoders bpe.py:134-158 /
 * wordpiece.py:160-179 when naive != 0).  The swt_small object owns a pinned, device-mapped input and output buffer, device
 * scratch and a stream; a call copies the text into the pinned buffer, launches ONE single-CTA kernel (pre-tokenizer +
 * encoder; no memo, no batching) and synchronises once.  Nothing is allocated per call.  Texts of up to
 * swt_small_max_bytes() bytes; *ids points into the object's output buffer (u32 ids) and stays valid until the next call.
 * which / table / pretok as in swt_tokenize_text_host.  n_words / h6_events may be NULL.  One call at a time per swt_small object
 * (the reference's classes are single-threaded); use one object per host thread otherwise. */
typedef struct swt_small swt_small;
int swt_small_create(int device, swt_small **out);
void swt_small_destroy(swt_small *s);
uint32_t swt_small_max_bytes(void);
int swt_tokenize_small(swt_small *s, const swt_pretok *pretok, int which, const void *table, int naive, const uint8_t *text,
                       uint32_t n_bytes, const uint32_t **ids, uint32_t *n_tokens, uint32_t *n_words, uint32_t *h6_events);
/* The same call with the (pre-tokenizer, table) pair bound once (hosts where every argument conversion counts, e.g. ctypes):
 * swt_tokenize_small_bound takes the text only; the results are read from the output buffer swt_small_output(s):
 * [0] status, [1] tokens, [2] H6 events, [3] words, u32 ids from word 8.  At most 64 bindings at a time (SWT_ERR_CAPACITY
 * beyond: use swt_tokenize_small); swt_small_unbind releases one, e.g. before its table is destroyed. */
int swt_small_bind(swt_small *s, const swt_pretok *pretok, int which, const void *table, int naive, uint32_t *binding);
void swt_small_unbind(swt_small *s, uint32_t binding);
const uint32_t *swt_small_output(const swt_small *s);
int swt_tokenize_small_bound(swt_small *s, uint32_t binding, const uint8_t *text, uint32_t n_bytes);
/* pinned host allocation helpers so integrators can give the pipeline DMA-able buffers */
int swt_host_alloc(void **ptr, size_t bytes);
void swt_host_free(void *ptr);

/* ---- HP-3: BPE training ---------------------------------------------------------------------------- */
/*
 * Replaces the merge loop of NaiveBPE.train (bpe.py:88-111): pair-frequency count (:90-95),
 * argmax with first-inserted tie-break (:102), vocabulary/merge-list update (:103-104) and the
 * merge-apply over every word (:108-111 -> _replace_pair :25-48).
 *
 * Inputs are the word types in first-occurrence order (bpe.py:77-81) as sequences of alphabet
 * ids 0..n_alpha-1 (one per distinct code point) with their frequencies.  Merged symbols get ids
 * n_alpha, n_alpha+1, ... in order of first creation of each DISTINCT string (a merge whose
 * concatenation already exists reuses that id and does not grow the vocabulary, bpe.py:103).
 *
 * Multi-GPU: word types are sharded by contiguous ranges; every rank holds `n_types_local`
 * types starting at global symbol slot `slot_base` and a replica of the pair-count table.
 * One training step is split in three phases so that the caller can run the two small
 * collectives (NCCL through torch.distributed) between them on the same stream:
 *
 *     swt_bpe_train_select   argmax over the replicated table + local tie-break scan
 *                            -> candidate record (16 B) at cand_ptr          [all_gather if world>1]
 *     swt_bpe_train_merge    resolve winner, name the new symbol, mark + apply the merge
 *                            in place, emit pair-count deltas at delta_ptr   [all_reduce(SUM) if world>1]
 *     swt_bpe_train_update   fold the deltas into the replicated table
 *
 * swt_bpe_train_steps runs `n_steps` whole steps back to back for world == 1.
 * All phases are asynchronous and self-gating: once the trainer halts (vocabulary reached,
 * no pairs left, table must grow, record buffer full, error) the remaining launches are no-ops.
 */
typedef struct swt_bpe_trainer swt_bpe_trainer;

/* SWT_TRAIN_WP runs the same machinery for NaiveWP.train (wordpiece.py:29-103): symbols
 * 0..n_alpha-1 are the initial symbol STRINGS ("a", "##a", ...; d_init_cps/d_init_off), the merged token is a + b[2:]
 * (:95) and each step takes the pair with the highest score pair_freq / (freq[a]*freq[b]) (:84-92, first inserted pair
 * on ties).  Single rank only; symbol-frequency products must stay below 2^53. */
#define SWT_TRAIN_BPE 0
#define SWT_TRAIN_WP 1

typedef struct swt_bpe_train_config {
    uint64_t n_types_local;    /* word types on this rank */
    uint64_t n_slots_local;    /* total symbols of those types (< 2^32) */
    uint64_t slot_base;        /* global index of this rank's first symbol slot */
    uint32_t n_alpha;          /* alphabet size (global) */
    int64_t  max_vocab;        /* stop when the vocabulary reaches this size (bpe.py:88) */
    int64_t  initial_vocab;    /* number of distinct code points present globally (bpe.py:75) */
    uint32_t max_word_len;     /* longest word type in symbols (global) */
    uint32_t record_cap;       /* merges recorded between two swt_bpe_train_read calls */
    uint32_t world_size;       /* number of ranks sharing the job */
    uint32_t rank;
    uint64_t table_cap;        /* pair-table slots (power of two); 0 = choose */
    uint32_t mode;             /* SWT_TRAIN_BPE or SWT_TRAIN_WP */
} swt_bpe_train_config;

/* host-visible snapshot of the trainer state */
typedef struct swt_bpe_train_state {
    uint32_t halt;             /* 0 running, 1 done (vocab reached), 2 done (no pairs), 3 table must grow,
                                  4 record buffer full, >=16 error */
    uint32_t n_recorded;       /* merges in the record buffer */
    uint64_t n_merges_total;   /* merges since create */
    int64_t  vocab_size;
    uint64_t n_symbols;        /* distinct symbol strings so far */
    uint64_t n_table_entries;
    uint64_t table_cap;
    uint64_t n_live_slots;     /* live symbols on this rank */
    uint64_t n_tie_steps;      /* steps in which several pairs attained the maximum (first-occurrence scan) */
    uint64_t n_tie_listed;     /* ... of which the tied pairs were few enough to be listed (filtered scan, no table probes) */
    uint64_t n_peer_barriers;  /* peer exchange: cross-GPU barriers passed ... */
    uint64_t peer_wait_cycles; /* ... and SM cycles one lane waited in them (rank skew + NVLink latency) */
    uint64_t peer_kernel_cycles[3]; /* diagnostic: SM cycles inside the candidate exchange, the delta exchange, thread 0 of the inbox accumulation */
} swt_bpe_train_state;

size_t swt_bpe_train_workspace_bytes(const swt_bpe_train_config *cfg);
/* d_syms/d_off/d_freq: this rank's types (u32 symbol ids, u64 offsets n_types_local+1, i64 freqs) */
int swt_bpe_train_create(const swt_bpe_train_config *cfg, const uint32_t *d_syms, const uint64_t *d_off,
                         const int64_t *d_freq, const uint32_t *d_init_cps /* WP only, else NULL */,
                         const uint64_t *d_init_off /* WP only: n_alpha+1 */, void *d_workspace, size_t workspace_bytes,
                         void *stream, swt_bpe_trainer **out);
void swt_bpe_train_destroy(swt_bpe_trainer *t);
/* device addresses/sizes of the exchange buffers, for the caller's collectives */
int swt_bpe_train_buffers(const swt_bpe_trainer *t, void **init_counts_ptr, uint64_t *init_counts_elems /* i64 */,
                          void **cand_ptr /* 2 x u64 */, void **cand_gather_ptr /* world x 2 x u64 */,
                          void **delta_ptr, uint64_t *delta_elems /* i64 */);
/* initial pair count (bpe.py:90-95 before the first merge).
 * n_alpha <= 4096: local dense count -> [all_reduce(SUM) over init_counts] -> table build.
 * larger alphabets (init_counts_elems == 0): count_local inserts this rank's counts straight into the pair table and
 * build_table is a no-op; sharded callers then swt_bpe_train_export_pairs (2 x u64 per entry: key, count; *d_n_out = number
 * of entries, which may exceed out_cap_entries -- call again with a larger buffer), all-gather the lists and
 * swt_bpe_train_import_pairs the lists of the OTHER ranks, so that every replica holds the global counts. */
int swt_bpe_train_count_local(swt_bpe_trainer *t, void *stream);
int swt_bpe_train_build_table(swt_bpe_trainer *t, void *stream);
int swt_bpe_train_export_pairs(swt_bpe_trainer *t, uint64_t *d_out, uint64_t out_cap_entries, uint64_t *d_n_out, void *stream);
int swt_bpe_train_import_pairs(swt_bpe_trainer *t, const uint64_t *d_in, uint64_t n_entries, void *stream);
int swt_bpe_train_select(swt_bpe_trainer *t, void *stream);
int swt_bpe_train_merge(swt_bpe_trainer *t, void *stream);
int swt_bpe_train_update(swt_bpe_trainer *t, void *stream);
int swt_bpe_train_steps(swt_bpe_trainer *t, uint32_t n_steps, void *stream);
/* Peer-memory exchange for the sharded trainer (world_size 2..8 GPUs of one box, NVLink / NVSwitch): instead of the caller's two
 * collectives per step, every rank pushes its tie-break candidate (tie steps only) and its list of touched (symbol, delta) pairs
 * straight into an inbox of every rank with P2P stores, followed by a flag barrier -- two small kernels inside the step, which
 * swt_bpe_train_steps then captures into its CUDA graph like the single-rank loop.  Every rank allocates one buffer of
 * swt_bpe_train_peer_bytes(cfg) bytes that ALL ranks can address (symmetric memory / CUDA IPC), zeroes it, and passes the
 * world_size device pointers (index = rank) to swt_bpe_train_set_peers after create and before the first step.  The initial
 * pair counts still go through the caller's all-reduce (once). */
size_t swt_bpe_train_peer_bytes(const swt_bpe_train_config *cfg);
int swt_bpe_train_set_peers(swt_bpe_trainer *t, void *const *peer_buffers, uint32_t n_peers);
int swt_bpe_train_exchange_candidates(swt_bpe_trainer *t, void *stream);
int swt_bpe_train_exchange_deltas(swt_bpe_trainer *t, void *stream);
/* diagnostic: n_rounds cross-GPU barriers and nothing else (every rank must call it with the same n_rounds) */
int swt_bpe_train_exchange_probe(swt_bpe_trainer *t, uint32_t n_rounds, void *stream);
/* synchronises `stream`, copies out the recorded merges (left,right,new ids + chosen pair count)
   and the state; resets the record buffer and clears halt==4. Arrays need record_cap entries. */
int swt_bpe_train_read(swt_bpe_trainer *t, uint32_t *h_left, uint32_t *h_right, uint32_t *h_new, int64_t *h_count,
                       swt_bpe_train_state *state, void *stream);
/* replaces the pair table by one of `new_cap` slots inside a caller-provided buffer; clears halt==3 */
size_t swt_bpe_train_table_bytes(uint64_t cap);
int swt_bpe_train_grow_table(swt_bpe_trainer *t, void *d_new_table, uint64_t new_cap, void *stream);
/* Maintenance between batches of steps (call it where the merges are read back, every few hundred steps; n_live_slots from the
 * state): rebuilds the per-chunk pair filter of the mark scan from the live pairs and, once fewer than 60 % of the slots are
 * live, compacts the word table so that the scans shrink with the corpus.  *kernels_changed = 1 when the kernels' extents
 * changed: CUDA graphs the CALLER captured over select / merge / update must be re-captured.  May synchronise and allocate. */
int swt_bpe_train_maintain(swt_bpe_trainer *t, uint64_t n_live_slots, int *kernels_changed, void *stream);
/* copies the current segmentation of this rank's types to the host: symbol ids (type w at its original offset) + per-type lengths */
int swt_bpe_train_read_corpus(swt_bpe_trainer *t, uint32_t *h_syms /* n_slots_local */, uint32_t *h_len /* n_types_local */,
                              void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SWT_H */
