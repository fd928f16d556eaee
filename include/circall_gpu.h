/* circall_gpu.h — C ABI of libcircall_gpu.so, the B200 (sm_100a) implementation of Circall's
 * quasi-mapping hot path (Circall_wt + Circall_bsj).
 *
 * Circall itself has no plugin / FFI layer: its mappers are stand-alone executables glued by files
 * (SURVEY.md §8 b).  Each entry point below therefore cites the piece of the reference's C++ that a
 * maintainer would replace with the call (paths relative to Circall_v1.0.1/ in datngu/Circall);
 * INTEGRATION.md shows the patch.  Plain pointers and sizes only; no exceptions cross the boundary;
 * every function returns 0 on success or a negative cg_status, with text in cg_last_error().
 * There is no CPU fallback: without a CUDA device every compute entry point fails with CG_ERR_CUDA.
 */
#ifndef CIRCALL_GPU_H
#define CIRCALL_GPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    CG_OK = 0,
    CG_ERR_ARG = -1,        /* bad argument */
    CG_ERR_CUDA = -2,       /* CUDA runtime error / no device */
    CG_ERR_IO = -3,         /* index directory unreadable / malformed */
    CG_ERR_BASE = -4,       /* a read holds a character outside ACGTNacgtn */
    CG_ERR_LENGTH = -5,     /* a read is longer than CG_MAX_READ_LEN or the text exceeds 2^31-1 */
    CG_ERR_CAPACITY = -6,   /* an internal fixed capacity was exceeded (reported, never dropped) */
    CG_ERR_STATE = -7       /* call sequence error (fetch without submit, ...) */
} cg_status;

#define CG_MAX_READ_LEN 256

typedef struct cg_index cg_index;       /* device-resident quasi-index */
typedef struct cg_session cg_session;   /* one per host thread per GPU */

/* message of the last failing call on this thread */
const char* cg_last_error(void);
/* number of visible CUDA devices (0 when there is none) */
int cg_device_count(int* n);

/* ---- index -----------------------------------------------------------------------------------
 * cg_index_build: replaces SailfishIndex::build -> RapMap indexTranscriptsSA behind
 *   src/TxIndexer.cpp:202-227.  `text` is the concatenation of the (already normalised, upper-case
 *   ACGT) transcripts, each followed by '$'.  Builds on `device`: 2-bit text blocks, suffix array
 *   (GPU prefix doubling), k-mer table.  k must be odd, <= 31 (TxIndexer.cpp:87,208-214).
 * cg_index_build_fasta: same from a FASTA file, applying RapMap's record normalisation
 *    and keeping names up to the first whitespace.
 * cg_index_save / cg_index_open: replace the index directory I/O of SailfishIndex::load
 *   (include/ReadExperiment.hpp:66-67).  The directory holds the file set RapMap's indexer writes behind TxIndexer
 *   (header.json, versionInfo.json, txpInfo.bin, sa.bin, rsd.bin — SURVEY.md B7: recollected, unpinned; hash.bin, a sparsehash
 *   dump, is neither read nor written) plus this library's cache (cg_meta.bin, cg_text.bin, cg_hash.bin).  cg_index_open takes a
 *   directory with the cache as it is and a directory without one — an existing -txIdx / -bsjIdx written by the reference's
 *   TxIndexer — by reading the text from txpInfo.bin and rebuilding suffix array and k-mer table on the device.
 * cg_index_dir_info: what a directory holds, read on the host (no device needed): text length, transcripts, k;
 *   layout 1 = the reference's file set only, 2 = with this library's cache.
 */
int cg_index_build(int device, const char* text, uint64_t n, int k, cg_index** out);
int cg_index_build_fasta(int device, const char* fasta_path, int k, cg_index** out);
int cg_index_save(const cg_index* idx, const char* dir);
int cg_index_open(const char* dir, int device, cg_index** out);
void cg_index_close(cg_index* idx);

typedef struct {
    uint64_t text_len;      /* characters including '$' */
    uint32_t n_tx;          /* transcripts (txpNames.size(), ReadExperiment.hpp:122) */
    uint64_t n_kmers;       /* distinct k-mers in the table */
    uint64_t hash_capacity; /* slots */
    int k;
    int device;
    uint64_t device_bytes;  /* HBM held by the index */
    double build_seconds;   /* 0 for an opened index */
} cg_index_info;
int cg_index_get_info(const cg_index* idx, cg_index_info* info);
int cg_index_dir_info(const char* dir, cg_index_info* info, int* layout);
/* txpNames[t], txpLens[t], txpOffsets[t]  (ReadExperiment.hpp:126-135) */
const char* cg_index_tx_name(const cg_index* idx, uint32_t t);
int cg_index_set_tx_name(cg_index* idx, uint32_t t, const char* name);
uint32_t cg_index_tx_len(const cg_index* idx, uint32_t t);
uint32_t cg_index_tx_offset(const cg_index* idx, uint32_t t);
/* copy the suffix array to host memory (n int32) — used by the parity tests and by bench.py to hand
 * the array to the CPU baseline, which verifies it before use */
int cg_index_copy_sa(const cg_index* idx, int32_t* out);

/* ---- collector (one call = SACollectorCircall::operator() on every read of the batch) ----------
 * Mirrors include/SACollectorCircall.hpp:24-31: for each read returns the post-vote fwdSAInts /
 * rcSAInts and the QuasiAlignment hit list.  This is the row-A3 parity surface; the pipeline below
 * does not go through host-visible intervals.
 *   bases/offsets : concatenated ASCII reads, n_reads+1 offsets
 *   ints_*        : 4 int32 per interval (begin, end, len, queryPos); ints_off has n_reads+1 entries
 *   hits_*        : tid / pos / fwd per hit, hits_off has n_reads+1 entries
 * Buffers are owned by the library until the next cg_collect call on this thread or cg_collect_free.
 */
typedef struct {
    uint64_t n_reads;
    const uint8_t* found;                 /* operator() return value per read */
    const uint64_t* fwd_off; const int32_t* fwd_ints;
    const uint64_t* rc_off;  const int32_t* rc_ints;
    const uint64_t* hits_off; const uint32_t* hit_tid; const int32_t* hit_pos; const uint8_t* hit_fwd;
} cg_collect_result;
int cg_collect(const cg_index* idx, const char* bases, const uint64_t* offsets, uint64_t n_reads,
               int strict_check, int lce_mode, cg_collect_result* out);
void cg_collect_free(void);

/* ---- mate-hit intersection ---------------------------------------------------------
 * rapmap::utils::mergeLeftRightHitsFuzzy as Circall_pseudo.cpp:271-280 runs it (maxNumHits = 200):
 * for every pair, two tid-sorted hit lists in, joint hits out (7 int32 each: tid, pos, fwd, matePos,
 * mateIsFwd, fragLen, mateStatus[3 = paired, 1 = left orphan, 2 = right orphan]).
 */
int cg_merge_pairs(int device, uint64_t n_pairs, const uint64_t* left_off, const uint32_t* left_tid,
                   const int32_t* left_pos, const uint8_t* left_fwd, const uint8_t* left_seeded,
                   const uint64_t* right_off, const uint32_t* right_tid, const int32_t* right_pos,
                   const uint8_t* right_fwd, const uint8_t* right_seeded, uint32_t read_len,
                   uint32_t max_num_hits, uint64_t* joint_off /* n_pairs+1 */, int32_t* joint /* cap*7 */,
                   uint64_t joint_cap, uint8_t* too_many /* n_pairs */);

/* ---- pipeline session ---------------------------------------------------------------------------
 * One session = the per-pair bodies of Circall_wt (src/Circall_wt.cpp:237-660) and Circall_bsj
 * (src/Circall_bsj.cpp:225-505) fused: both mates are mapped to tx, every (pair, unmapped-but-seeded
 * mate) record is mapped to bsj, span-filtered and counted, without the intermediate
 * UN_read{1,2}_*.fa files of Circall_source.sh:273-294.
 */
typedef struct {
    int lce_mode;                /* 0 = LCE_LITERAL (default), 1 = LCE_CLEAN   */
    uint32_t max_batch_pairs;    /* capacity hint; 0 = 1<<20 */
    int stage;                   /* 0 = wt then bsj (fused), 1 = wt only (bsj index may be NULL),
                                    2 = bsj only: mate 1 of every submitted pair is the record's read 1 */
    int keep_mate_detail;        /* also return per-mate flags / hit counts (parity tests) */
    int dump_intervals;          /* debug surface of row A3: also return, for every collector call of the batch, the post-vote
                                    fwdSAInts / rcSAInts (SACollectorCircall.hpp:442-490) as the tier that answered the read
                                    computed them, and which tier that was (parity tests; costs time and memory) */
} cg_opts;

typedef struct { uint32_t rec; uint32_t dir; uint32_t query_pos; uint32_t len; int32_t hit_pos; uint32_t tid; } cg_junction;

typedef struct {
    uint64_t n_pairs;
    /* Circall_wt export list, in input order (Circall_wt.cpp:350-365,525-540):
     * code 1 = mate 1 unmapped-but-seeded (headers ___11/___22, mate 1 first);
     * code 2 = mate 2 (headers ___21/___12, mate 2 first).  A pair can appear twice. */
    uint64_t n_export; const uint32_t* exp_pair; const uint8_t* exp_code;
    /* Circall_bsj junction table rows (Circall_bsj.cpp:339-341,366-368); rec indexes the export list */
    uint64_t n_junc; const cg_junction* junc;
    /* records written to BSJ_read{1,2}_*.fa (Circall_bsj.cpp:390-399); rec indexes the export list */
    uint64_t n_cand; const uint32_t* cand_rec;
    /* multi-transcript equivalence classes of this batch (EquivalenceClassBuilder.hpp:112-130):
     * class c covers eq_tids[eq_off[c] .. eq_off[c+1]) and was seen once per listed entry */
    uint64_t n_eq; const uint64_t* eq_off; const uint32_t* eq_tids;
    /* per mate (2 per pair): bit0 seeded, bit1 some interval has len == readLen; hits after the >200 clear */
    const uint8_t* mate_flags; const uint16_t* mate_nhits;
    /* per export record: jointHits.size() before the >200 clear; isSplit */
    const uint32_t* rec_nhits; const uint8_t* rec_split;
    double device_ms;            /* kernels of this batch, CUDA events on the session's stream */
    double kernel_ms[4];         /* of which: read packing, wt collector, wt commit + export compaction, bsj collector + span filter */
    uint64_t n_fallback_wt, n_fallback_bsj;   /* mates / records that needed the large-capacity general kernels */
    uint64_t n_tier2_wt, n_tier2_bsj;         /* mates / records the thread-per-read first tier handed to the warp-per-read kernels */
    /* cg_opts::dump_intervals (fetch with want_lists): per Circall_wt mate (2 per pair) / per Circall_bsj record
     * head = nF | nR << 12 | tier << 24 | found << 28 (tier 1 clean-match kernel, 2 / 3 lockstep thread tier lean / two-strand,
     * 4 / 5 warp kernels fast / general), off = index of the item's first interval in *_iv (4 int32 each: begin, end, len,
     * queryPos; the nF forward ones first) */
    const uint32_t* wt_iv_head; const uint32_t* wt_iv_off; const int32_t* wt_iv; uint64_t n_wt_iv;
    const uint32_t* bsj_iv_head; const uint32_t* bsj_iv_off; const int32_t* bsj_iv; uint64_t n_bsj_iv;
} cg_batch_result;

/* scalar counters (ReadExperiment.hpp:265-270) for the two tools */
typedef struct {
    uint64_t wt_num_observed, wt_num_mapped, wt_num_hits, wt_upper_bound;
    uint64_t bsj_num_observed, bsj_num_mapped, bsj_num_hits, bsj_upper_bound;
    uint32_t wt_read_len, bsj_read_len;   /* first read length seen (Circall_wt.cpp:303, Circall_bsj.cpp:310) */
} cg_counters;

/* Several sessions may be in flight on one device (each has its own stream: copies of one overlap kernels of the others).  The
 * kernels read the index views from per-device constant memory; sessions on the same indices share them, and a submit from a
 * session on OTHER indices first waits (on the device, stream-ordered) for the batches in flight on the current ones — correct for
 * any mix of sessions, concurrent only among sessions of one index pair. */
int cg_session_create(const cg_index* tx, const cg_index* bsj, const cg_opts* opts, cg_session** out);
void cg_session_destroy(cg_session* s);
/* host buffers (pageable or pinned): concatenated ASCII bases + n_pairs+1 offsets per mate file
 * (the jobs pair_sequence_parser hands to processReadsQuasi, Circall_wt.cpp:237-241) */
int cg_session_submit(cg_session* s, const char* r1, const uint64_t* off1, const char* r2, const uint64_t* off2,
                      uint64_t n_pairs);
/* same for a batch whose reads all have one length (what a sequencer writes): r1 / r2 hold n_pairs * read_len characters and
 * the offsets are made on the device, so they do not cross PCIe (16 of the 216 bytes per 2x100 pair) and are not checked on
 * the host.  Results are those of cg_session_submit with offsets i * read_len. */
int cg_session_submit_fixed(cg_session* s, const char* r1, const char* r2, uint64_t n_pairs, uint32_t read_len);
/* same, buffers already resident in this session's GPU memory (bench.py `value` leg); max_read_len bounds
 * every read of the batch (a longer read fails the batch with CG_ERR_LENGTH at fetch) */
int cg_session_submit_device(cg_session* s, const char* d_r1, const uint64_t* d_off1, const char* d_r2,
                             const uint64_t* d_off2, uint64_t n_pairs, uint32_t max_read_len);
/* wait for the batch and expose its results (host memory owned by the session until the next submit).
 * want_lists = 0 skips the device->host copy of the per-record lists (counters still accumulate). */
int cg_session_fetch(cg_session* s, int want_lists, cg_batch_result* out);
/* totals since session creation; per_bsj (may be NULL) receives the single-transcript equivalence-class
 * counts, one uint64 per BSJ reference (the `Count` column of rawCount.txt rows with Weight 1) */
int cg_session_counters(cg_session* s, cg_counters* c, uint64_t* per_bsj, uint64_t n_bsj);
/* device pointer of the dense accumulator [n_bsj per-BSJ counts | 8 scalar counters] as uint64, for the
 * single end-of-run all-reduce across GPUs (SURVEY.md §8 e); the caller owns the communicator */
int cg_session_accumulator(cg_session* s, void** dptr, uint64_t* n_u64);
/* The one collective of the path (the reference has none: its workers share one EquivalenceClassBuilder and atomics,
 * src/Circall_bsj.cpp:402-491): all-reduce (sum) of the accumulators of `n` sessions of ONE process, on one or several devices —
 * sessions of a device are added up on it, the per-device totals go through one ncclAllReduce (uint64, over NVLink; NCCL is bound at
 * first use) and every session ends up holding the grand total, so that cg_session_counters on any of them returns the run's
 * totals.  Call once, after the last fetch.  (Ranks of a multi-process job reduce cg_session_accumulator's vector through their own
 * communicator instead: bench.py does with torch.distributed.) */
int cg_sessions_allreduce(cg_session* const* sessions, int n);
/* kernels launched by this session so far (bench.py gpu_launches) */
uint64_t cg_session_launches(const cg_session* s);
/* device / pinned-host allocation helpers (resident-input leg, pinned staging buffers of the host streamer) */
int cg_device_alloc(int device, uint64_t bytes, void** dptr);
int cg_device_free(int device, void* dptr);
int cg_device_upload(int device, void* dptr, const void* src, uint64_t bytes);
int cg_device_download(int device, void* dst, const void* dptr, uint64_t bytes);
int cg_host_alloc(uint64_t bytes, void** ptr);
int cg_host_free(void* ptr);
/* overwrite a buffer larger than the L2 so that the next timed pass starts cold */
int cg_device_flush_l2(int device);

/* random 32-B-sector gather microbenchmark (SURVEY.md §8 d "random-sector ceiling"): every thread
 * issues `per_thread` dependent-free 32-B loads at hashed addresses inside a `bytes`-sized buffer.
 * Returns achieved GB/s of useful sectors. */
int cg_bench_gather(int device, uint64_t bytes, uint32_t per_thread, int iters, double* gbps);
/* the same with another load instruction (mode 0 = cg_bench_gather's ld.global.nc; 1-3 L2::64B/128B/256B prefetch size, 4 L1::no_allocate,
 * 5 .cg, 6 plain ld.global, 7 no_allocate + L2::64B, 8 .cv, 9 both 16-byte halves of the sector): what one missing sector costs in DRAM
 * bytes depends on it, and that cost is the ceiling of the whole random-access path */
int cg_bench_gather_mode(int device, uint64_t bytes, uint32_t per_thread, int iters, int mode, double* gbps);

/* The read-packing kernel (K1) on its own, host buffers in and out: what a session does to a batch before the
 * collectors run (2-bit codes as the collector reads them, SACollectorCircall.hpp:117-128,243-260: base j of a
 * 32-base word at bits 2j, A 0 C 1 G 2 T 3, any other character code 0 with its bit set in the N mask).
 * packed: 2 * n_pairs records (mate 1 then mate 2 of each pair) of W2 + WM 64-bit words, W2 = ceil(max_len / 32),
 * WM = ceil(W2 / 2); lens: 2 * n_pairs; status bit 0: a character outside ACGTNacgtn, bit 1: a read longer than
 * max_len (cut).  general != 0 runs the every-character formulation instead of the all-bases short cut (the two must
 * agree bit for bit: tests/test_gpu_pipeline.py).  iters timed launches after one warm-up; ms_per_launch may be NULL. */
int cg_pack_reads(int device, const char* r1, const uint64_t* off1, const char* r2, const uint64_t* off2, uint64_t n_pairs,
                  int max_len, int general, int iters, uint64_t* packed, uint16_t* lens, int* status, double* ms_per_launch);

/* The index builder's sort / scan / selection primitives on caller-provided host data (cg_sortscan.cuh; test surface of row N2:
 * the reference's indexer sorts with libdivsufsort, src/TxIndexer.cpp:202-227 -> SailfishIndex::build [U]).
 * cg_debug_sort_pairs: stable sort of (keys[i], vals[i]) by key bits [begin_bit, end_bit), in place.
 * cg_debug_scan: mode 0 exclusive sum, 1 inclusive max of data[0..n) in place; 2: data holds flags, out_idx receives the indices of
 * the non-zero ones in ascending order, *n_out their number. */
int cg_debug_sort_pairs(int device, uint64_t* keys, uint32_t* vals, uint64_t n, int begin_bit, int end_bit);
int cg_debug_scan(int device, uint32_t* data, uint64_t n, int mode, uint32_t* out_idx, uint64_t* n_out);

#ifdef __cplusplus
}
#endif
#endif
