/* muscato_b200.h -- C ABI of the B200-native screen -> group -> confirm -> combine hot path.
 *
 * The reference (kshedden/muscato) has no FFI: its boundary is process level
 * (executables + config.json + files in TempDir).  This header is the thin layer a
 * cgo shim or the stage-compatible C++ executables bind.  Each
 * entry point names the reference code it replaces (file:line under the reference
 * tree).  Plain C, opaque handle, caller-owned inputs (borrowed for the call),
 * library-owned outputs released with msc_free, int return codes (0 = ok), no
 * exceptions across the boundary.  A context is bound to one CUDA device and is not
 * thread-safe (callers serialise).  There is NO CPU fallback: without a usable
 * sm_100 device msc_create fails with MSC_ERR_CUDA.
 */
#ifndef MUSCATO_B200_H
#define MUSCATO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSC_MAX_WINDOWS 32
#define MSC_MAX_WINDOW_WIDTH 50   /* wider windows are undefined in the reference (100 - q2 < W at target position 0) */
#define MSC_MAX_READ_LENGTH 1024

enum {
  MSC_OK = 0,
  MSC_ERR_CONFIG = 1,   /* invalid / unsupported configuration */
  MSC_ERR_INPUT = 2,    /* malformed input buffers */
  MSC_ERR_CUDA = 3,     /* CUDA runtime failure (message in msc_last_error) */
  MSC_ERR_STATE = 4,    /* call order violated (e.g. screen before set_reads) */
  MSC_ERR_NOMEM = 5,
  MSC_ERR_IO = 6,
  MSC_ERR_AGAIN = 7     /* a deferred run has to be repeated (see MSC_STAGE_DEFER) */
};

enum { MSC_MATCH_FIRST = 0, MSC_MATCH_BEST = 1 };

/* The hot-path subset of utils.Config (utils/config.go:10-101) plus device selection.
 * BloomSize / NumHash are accepted by the stage executables and ignored: the
 * rolling-hash Bloom screen (cmd/muscato_screen/main.go:86-253) is replaced by an
 * exact key table behind a blocked Bloom front. */
typedef struct msc_config {
  int32_t n_windows;                  /* len(Config.Windows), 1..MSC_MAX_WINDOWS */
  int32_t windows[MSC_MAX_WINDOWS];   /* Config.Windows (left end of each window) */
  int32_t window_width;               /* Config.WindowWidth, 1..MSC_MAX_WINDOW_WIDTH */
  int32_t max_read_length;            /* Config.MaxReadLength, <= MSC_MAX_READ_LENGTH */
  double  pmatch;                     /* Config.PMatch; nmiss = int((1-PMatch)*float64(L)), confirm main.go:198 */
  int32_t min_dinuc;                  /* Config.MinDinuc (utils/entropy.go) */
  int32_t mmtol;                      /* Config.MMTol (cmd/muscato_combine_windows/main.go:36-60) */
  int64_t max_matches;                /* Config.MaxMatches (cmd/muscato_confirm/main.go:233-242, 424-448) */
  int32_t match_mode;                 /* MSC_MATCH_FIRST / MSC_MATCH_BEST */
  int32_t device;                     /* CUDA device ordinal */
  int32_t bloom_bits_per_key;         /* 0 = default; tuning only, never changes results */
  int32_t keep_ascii;                 /* keep device copies of the ASCII inputs (needed by msc_rebuild) */
  int32_t reserved[8];
} msc_config;

/* One confirmed alignment: the fields of an rmatch/matches.txt line
 * (cmd/muscato_confirm/main.go:221-230) as integers.  The text columns are
 * read = reads[read_id], target = targets[gene_id][pos : pos+len(read)]. */
typedef struct msc_match {
  uint32_t read_id;   /* index into the reads passed to msc_set_reads (reads_sorted order) */
  uint32_t gene_id;   /* 0-based target line index (cmd/muscato_screen/main.go:440-452) */
  uint32_t pos;       /* mpos - len(mlft): start of the read within the target */
  uint32_t nx;        /* mismatch count */
} msc_match;

typedef struct msc_stats {
  uint64_t n_reads, n_keys, n_key_groups, table_slots, bloom_bytes;
  uint64_t n_targets, target_bases, positions_probed;
  uint64_t n_candidates, n_pairs, n_pass, n_matches_pre, n_matches;
  uint64_t n_overflow_groups;         /* (window,key) groups that hit MaxMatches */
  uint64_t h2d_bytes, d2h_bytes;
  uint64_t kernel_launches;           /* kernels launched by this library since msc_reset_stats */
  float ms_pack_reads, ms_build, ms_pack_targets, ms_scan, ms_expand, ms_confirm, ms_combine;
  float ms_scan_kernel;               /* the scan kernel alone (CUDA events on the launch stream) */
  float reserved_f[8];                /* [0] positions that passed the Bloom front, [1] ms of msc_prep_reads on the device,
                                         [2] exact front: launches of the scan per step (slices of the 4^W-bit map), 0 = Bloom front */
} msc_stats;

typedef struct msc_ctx msc_ctx;

/* Library / build identification (no GPU needed).  msc_struct_size lets a binding verify
 * its mirror of the structs: which = 0 msc_config, 1 msc_match, 2 msc_stats, 3 msc_key_rec,
 * 4 msc_cand_rec. */
const char* msc_version(void);
uint64_t msc_struct_size(int which);

/* Create / destroy a context on config.device.  Validates the configuration like
 * checkArgs (cmd/muscato/main.go:833-904) does for the fields it owns.  On failure
 * returns NULL and, if err/errlen are given, writes a message there. */
msc_ctx* msc_create(const msc_config* config, char* err, uint64_t errlen);
void msc_destroy(msc_ctx* ctx);
const char* msc_last_error(const msc_ctx* ctx);

/* Unique reads in reads_sorted.txt order (first field of each line,
 * cmd/muscato_screen/main.go:165-172; cmd/muscato_window_reads/main.go:100-101):
 * read i is ascii[offs[i] .. offs[i+1]).  Alphabet A,C,G,T,X (any other byte is
 * treated as X).  Uploads, 2-bit packs on the device and builds the window-key
 * table: replaces buildBloom (cmd/muscato_screen/main.go:116-207), muscato_window_reads
 * (cmd/muscato_window_reads/main.go:94-141) and sortWindows (cmd/muscato/main.go:237-304).
 * n_reads * n_windows <= 2^30 per call; larger read sets are fed in batches (reads are independent
 * of each other: every rule of the path, MMTol included, is per read).
 * The upload is chunked (2^22 reads per chunk): a chunk's offsets are validated on the host, its
 * copy is queued, and its 2-bit pack + window pass run on the device under the next chunk's copy.
 * The call returns when the caller's buffers are no longer needed; the table build is stream
 * ordered behind it and its counters are booked by the next synchronising call. */
int msc_set_reads(msc_ctx* ctx, const uint8_t* ascii, const uint64_t* offs, uint64_t n_reads);

/* Same as msc_set_reads for buffers that are already in the memory of ctx's device (e.g. the
 * read set was uploaded once by rank 0 and broadcast over NVLink with NCCL).  offs has
 * n_reads + 1 entries ending at total_bytes; it is validated on the device. */
int msc_set_reads_device(msc_ctx* ctx, const uint8_t* d_ascii, const uint64_t* d_offs, uint64_t n_reads,
                         uint64_t total_bytes);

/* Device-side prepReads: replaces, for the sequence column,
 * `muscato_prep_reads | sort | muscato_uniqify` (cmd/muscato/main.go:152-221).  raw read i is
 * raw_ascii[raw_offs[i] .. raw_offs[i+1]) as it stands in the fastq file (any bytes, any length).
 * On the device: bytes outside A/C/G/T become X and reads are cut to MaxReadLength
 * (cmd/muscato_prep_reads/main.go:33-44, :67-69), reads shorter than min_read_length are skipped
 * (:59-62), the sequences are sorted bytewise (a proper prefix sorts first, as `LC_ALL=C sort` does
 * with the '\t' that follows the sequence) and equal sequences are collapsed
 * (cmd/muscato_uniqify/main.go:113-135).  The unique reads are installed as the context's read
 * set exactly as msc_set_reads would (pack + key table).  n_kept = reads that passed the length
 * filter, n_unique = distinct sequences. */
int msc_prep_reads(msc_ctx* ctx, const uint8_t* raw_ascii, const uint64_t* raw_offs, uint64_t n_raw,
                   int32_t min_read_length, uint64_t* n_kept, uint64_t* n_unique);

/* Grouping left by msc_prep_reads: perm[0..n_kept) = raw read indices in sorted order,
 * group_start[u]..group_start[u+1] = the members of unique read u (n_unique + 1 entries): count =
 * group size, names = the members' names (the host joins them in bytewise name order, which is
 * what sorting the `seq\tname` lines gives).  Buffers are caller-owned. */
int msc_fetch_read_groups(msc_ctx* ctx, uint32_t* perm, uint32_t* group_start);

/* The unique reads themselves (X-substituted, truncated), reads_sorted order: ascii needs
 * msc_unique_reads_bytes() bytes, offs n_unique + 1 entries. */
uint64_t msc_unique_reads_bytes(msc_ctx* ctx);
int msc_fetch_unique_reads(msc_ctx* ctx, uint8_t* ascii, uint64_t* offs);

/* Targets in GeneFileName order (gene id = index, cmd/muscato_screen/main.go:440-452):
 * target g is ascii[offs[g] .. offs[g+1]).  Total length < 2^32 - 4096 bases per call
 * (shard larger databases by target range).  Uploads and 2-bit packs on the device. */
int msc_set_targets(msc_ctx* ctx, const uint8_t* ascii, const uint64_t* offs, uint64_t n_targets);

/* Targets that are ALREADY 2-bit packed (the persistent target cache `<GeneFileName>.2bit` of the stage
 * executable, SURVEY.md 8(f) row f4: the database is parsed and packed once, later runs upload
 * 0.25 bytes per base instead of re-reading the text).  Layout of common.cuh: base i of the concatenated
 * stream of all targets sits in word i >> 5 at bits 2*(i & 31), codes A=0 C=1 T=2 G=3; xplane (same
 * spacing, bit 2*(i & 31) set = the base is X, its code bits 0) may be NULL when no target contains X.
 * words / xplane have (offs[n_targets] + 31) / 32 entries.  Same limits as msc_set_targets. */
int msc_set_targets_packed(msc_ctx* ctx, const uint64_t* words, const uint64_t* xplane, const uint64_t* offs,
                           uint64_t n_targets);

/* The targets of ANOTHER context on the same device, copied device to device (no host traffic).  Read parts are
 * independent of each other (every rule of the path is per read), so a second context can screen another part of
 * the reads against the same database while the first one is still uploading or computing: its read upload runs
 * under the other context's scan + confirm.  src must have targets; it may replace them once this call has returned. */
int msc_set_targets_from(msc_ctx* ctx, msc_ctx* src);

/* The packed form of the current targets (after msc_set_targets), to write such a cache: words and
 * xplane need msc_packed_target_words() entries each; *has_x receives whether any base is X. */
uint64_t msc_packed_target_words(const msc_ctx* ctx);
int msc_fetch_packed_targets(msc_ctx* ctx, uint64_t* words, uint64_t* xplane, int32_t* has_x);

/* Re-run the device-side pack/build from the resident ASCII copies (keep_ascii=1);
 * what: 1 = reads (+ key table), 2 = targets, 3 = both.  Used to time the hot path
 * with inputs already in HBM. */
int msc_rebuild(msc_ctx* ctx, int what);

/* Streaming scan of the packed targets against the key table, compacting
 * (key group, target position) candidates: replaces search/processSeq/checkWin/harvest
 * (cmd/muscato_screen/main.go:220-480). */
int msc_screen(msc_ctx* ctx);

/* Expand candidates x reads of the key group into pairs (replaces sortBloom,
 * cmd/muscato/main.go:318-385, and the merge join cmd/muscato_confirm/main.go:375-416),
 * count mismatches, apply the fit / mismatch rules and MaxMatches
 * (searchpairs, cmd/muscato_confirm/main.go:171-250, 424-448), de-duplicate across
 * windows (sort -u, cmd/muscato/main.go:453-463).  Leaves per-read best mismatch
 * counts in a device array (msc_best_device). */
int msc_confirm(msc_ctx* ctx);

/* Device pointer to uint32 best_nx[n_reads] (MSC_NO_MATCH = no match; positive as int32
 * too, so a signed MIN works) valid between msc_confirm and msc_combine.  With targets
 * sharded over several GPUs the host min-all-reduces this array across ranks (NCCL)
 * before msc_combine. */
#define MSC_NO_MATCH 0x7F7F7F7Fu
void* msc_best_device(msc_ctx* ctx);

/* Keep matches with nx <= best[read] + MMTol and group them by read:
 * replaces writebest (cmd/muscato_combine_windows/main.go:36-60). */
int msc_combine(msc_ctx* ctx);

/* Device pointer to the n surviving matches (msc_match layout, grouped by read_id) left by
 * msc_combine; valid until the next call that changes the context.  Lets the host gather
 * the compacted matches of all ranks over NCCL without a host round trip. */
void* msc_matches_device(msc_ctx* ctx, uint64_t* n);

/* Copy the surviving matches to the host, ordered by (read_id, gene_id, pos) (the order is
 * established on the device).  *out is owned by the library: release with msc_free. */
int msc_fetch_matches(msc_ctx* ctx, msc_match** out, uint64_t* n);

/* Same, into a caller-owned buffer of `capacity` records (e.g. pinned memory: no staging copy).
 * *n receives the number of matches; MSC_ERR_NOMEM if they do not fit. */
int msc_fetch_matches_into(msc_ctx* ctx, msc_match* dst, uint64_t capacity, uint64_t* n);

/* Reads without any confirmed match, ascending (= reads_sorted order, the order of the non-match
 * fastq): what muscato_nonmatch (cmd/muscato_nonmatch/main.go:57-113) derives from results.txt with
 * a Bloom filter, taken exactly from the per-read best array on the device.  Valid after
 * msc_confirm (with several target shards: after the MIN all-reduce of msc_best_device).
 * *n receives the count; MSC_ERR_NOMEM if it exceeds `capacity` (ids may be NULL to query *n). */
int msc_fetch_nonmatch(msc_ctx* ctx, uint32_t* ids, uint64_t capacity, uint64_t* n);

/* msc_screen + msc_confirm + msc_combine enqueued back to back on the context's stream with a
 * single host synchronisation at the end (intermediate counts stay on the device). */
int msc_run(msc_ctx* ctx);

/* msc_rebuild(what) followed by msc_run, again with a single synchronisation: one full pass of
 * the hot path over inputs that are already resident in HBM (keep_ascii=1). */
int msc_rebuild_and_run(msc_ctx* ctx, int what);

/* General form: optional rebuild (what = 0..3) followed by any prefix-consistent subset of the
 * stages, one synchronisation.  A multi-GPU host runs SCREEN|CONFIRM, all-reduces
 * msc_best_device over the ranks, then runs COMBINE. */
enum { MSC_STAGE_SCREEN = 1, MSC_STAGE_CONFIRM = 2, MSC_STAGE_COMBINE = 4, MSC_STAGE_DEFER = 8 };
int msc_run_stages(msc_ctx* ctx, int rebuild_what, int stages);

/* Stream-ordered multi-GPU step (no host round trip between the stages): with
 * SCREEN | CONFIRM | DEFER the stages are only ENQUEUED on msc_stream(ctx); the host orders its
 * NCCL MIN all-reduce of msc_best_device(ctx) after them on the same stream (or a stream that
 * waits on it) and then calls msc_run_stages(ctx, 0, MSC_STAGE_COMBINE), which enqueues the combine,
 * synchronises ONCE and completes all three stages.  If a bounded output buffer turned out too
 * small (it has been grown) or a key group exceeds MaxMatches, that call returns MSC_ERR_AGAIN and
 * the caller repeats the sequence (in the MaxMatches case without MSC_STAGE_DEFER). */
void* msc_stream(msc_ctx* ctx);   /* the cudaStream_t every kernel of the context runs on */

/* MaxMatches with the targets sharded by gene range over several contexts / ranks.
 * qinsert / "first" (cmd/muscato_confirm/main.go:233-242, :424-448) bound the result set of a
 * (window, k-mer) group over ALL targets, in the order of the globally sorted smatch_<k> file
 * (cmd/muscato/main.go:318-385), so a group whose passing pairs are spread over the shards cannot be
 * truncated by any one of them.  msc_set_shards(ctx, n) (n > 1) switches a context to the sharded
 * protocol:
 *   - msc_confirm flags key groups with more than MaxMatches / n passing pairs (if no shard has
 *     one, no group exceeds MaxMatches in total) and resolves nothing by itself; the flag is stored
 *     in element [n_reads] of the msc_best_device array, i.e. the host all-reduces n_reads + 1
 *     elements and every rank learns it with the exchange it performs anyway;
 *   - after msc_combine, msc_shard_overflow(ctx) says whether any shard flagged a group.  If so
 *     (rare path, host assisted): every rank calls msc_overflow_keys, the ranks exchange the key
 *     fingerprints (they identify a k-mer independently of the rank), every rank calls
 *     msc_divert_groups with the union -- which re-runs the pair kernel, keeps the undiverted
 *     matches and their per-read minimum in the context and returns the diverted passing pairs as
 *     records of msc_diverted_record_bytes(ctx) bytes (read, gene_base + gene, pos, nx, window, and
 *     the candidate's left / right context bytes, which the reference's candidate order compares) --
 *     one rank concatenates the records of all ranks and calls msc_replay_diverted, which replays
 *     the reference's sequential truncation and returns the surviving matches (global gene ids).
 *     The host folds their nx into the best array, all-reduces it again, calls msc_combine on every
 *     rank and merges the survivors that meet the MMTol rule with the gathered matches.
 * Outputs of msc_overflow_keys / msc_divert_groups / msc_replay_diverted are released with msc_free. */
int msc_set_shards(msc_ctx* ctx, int32_t n_shards);
int msc_shard_overflow(const msc_ctx* ctx);
int msc_overflow_keys(msc_ctx* ctx, uint64_t** keys, uint64_t* n);
uint32_t msc_diverted_record_bytes(const msc_ctx* ctx);
int msc_divert_groups(msc_ctx* ctx, const uint64_t* keys, uint64_t n_keys, uint32_t gene_base, uint8_t** recs, uint64_t* n_recs);
int msc_replay_diverted(msc_ctx* ctx, const uint8_t* recs, uint64_t n_recs, msc_match** out, uint64_t* n);

/* Per-stage timers (ms_pack_reads .. ms_combine) cost one event record between kernels per stage
 * boundary (~2.5 us each on a B200); on = 0 leaves them out -- ms_scan / ms_scan_kernel are always
 * measured.  Default: on (MSC_STAGE_EVENTS=0 in the environment turns them off at msc_create). */
int msc_set_stage_timing(msc_ctx* ctx, int on);

int msc_get_stats(const msc_ctx* ctx, msc_stats* out);
void msc_reset_stats(msc_ctx* ctx);
void msc_free(void* p);

/* Parity-debugging taps (SURVEY.md App. B checkpoints).  Each returns a library-owned
 * array released with msc_free.
 *   msc_dump_keys: every valid (window, read) key the table holds -- the content of
 *     win_<k>_sorted (cmd/muscato_window_reads/main.go:120-126) as (window, read_id) pairs.
 *   msc_dump_candidates: (gene_id, window start p, read_id, window) for every candidate x
 *     group member whose window k-mer matches exactly -- smatch_<k> joined with
 *     win_<k>_sorted on the k-mer (cmd/muscato_confirm/main.go:382-393), before the fit
 *     and mismatch rules. */
typedef struct msc_key_rec { uint32_t window; uint32_t read_id; } msc_key_rec;
typedef struct msc_cand_rec { uint32_t gene_id; uint32_t p; uint32_t read_id; uint32_t window; } msc_cand_rec;
int msc_dump_keys(msc_ctx* ctx, msc_key_rec** out, uint64_t* n);
int msc_dump_candidates(msc_ctx* ctx, msc_cand_rec** out, uint64_t* n);

#ifdef __cplusplus
}
#endif
#endif /* MUSCATO_B200_H */
