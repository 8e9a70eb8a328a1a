// api.cu -- context management and the C ABI of include/muscato_b200.h.
//
// One context = one CUDA device + one stream.  There is deliberately no CPU fallback: every
// entry point that computes anything launches the sm_100a kernels in this directory and fails
// with MSC_ERR_CUDA when that is impossible.
//
// Execution model: every stage is *enqueued* on the context's stream without a host round
// trip -- element counts produced by one kernel (candidates, pairs, matches) stay in a device
// counter block that the next kernels read, and launches use persistent / grid-stride grids
// whose shape does not depend on those counts.  A call synchronises once, at its end, to read
// the counter block (one 128-byte D2H) and to check that the bounded output buffers were large
// enough; if one was not, it is grown and the enqueue is repeated.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/muscato_b200.h"
#include "build.cuh"
#include "combine.cuh"
#include "common.cuh"
#include "confirm.cuh"
#include "pack.cuh"
#include "prefix.cuh"
#include "readprep.cuh"
#include "scan.cuh"
#include "scan_direct.cuh"

using namespace msc;

namespace {

// Kernel launch with the programmatic-stream-serialization attribute (PDL): the kernel may be
// launched while its predecessor in the stream is still running; every kernel starts with
// pdl_enter() (common.cuh), which waits for the predecessor's completion before touching memory.
template <typename... KArgs, typename... Args>
cudaError_t launch_k(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                     Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1u : 0u;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    bytes += 64;  // slack for 16-byte rounded fills and read-past-the-end word loads
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      want = bytes;
      e = cudaMalloc(&p, want);
    }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const { return reinterpret_cast<T*>(p); }
};
// A DevBuf that is a local of a host function: freed on every exit path.
struct ScopedDevBuf : DevBuf {
  ScopedDevBuf() = default;
  ScopedDevBuf(const ScopedDevBuf&) = delete;
  ScopedDevBuf& operator=(const ScopedDevBuf&) = delete;
  ~ScopedDevBuf() { release(); }
};

// Device counter block (unsigned long long each).
enum Counter {
  C_NKEYS = 0, C_NGROUPS, C_NDUP, C_SCRATCH, C_NCAND, C_BLOOMPASS, C_NMATCH, C_NPASS, C_NOVER, C_NOUT, C_NPAIRS,
  C_NDUMMY, C_NLONG, C_PAD1, C_TGX,  // C_NLONG, C_TGX at even indices; C_TGX = "some target word has X"
  C_PAD2, C_PREP_KEPT, C_PREP_UNIQ, C_PREP_BYTES, C_PAD3,  // C_PREP_KEPT at an even index (16)
  C_PREP_TIE, C_SHARDOVER,                                 // C_PREP_TIE (20): longest unsorted run of prep_tiefix_kernel;
                                                           // C_SHARDOVER: some target shard flagged a MaxMatches candidate group
  C_COUNT = 24  // groups that are cleared together start at even indices (16-byte aligned)
};

// Stage boundary events.
enum Ev {
  EV_PACKR0 = 0, EV_PACKR1, EV_BUILD1, EV_PACKT0, EV_PACKT1, EV_SCAN0, EV_SCAN1, EV_EXPAND1, EV_CONFIRM1, EV_COMB0,
  EV_COMB1, EV_COUNT
};

}  // namespace

struct msc_ctx {
  msc_config cfg{};
  WinCfg win{};
  int device = 0;
  int sm_count = 148;
  int scan_grid = 0, confirm_grid = 0;
  bool scatter_attr_set = false;
  const void* scan_fn_sized = nullptr;  // the scan kernel instance scan_grid was computed for
  cudaStream_t stream = nullptr;       // every kernel runs here
  cudaStream_t copy_stream = nullptr;  // input H2D copies: overlap with kernels of the previous input
  cudaEvent_t ev_copy = nullptr;
  // "the kernels that read the upload destinations have been enqueued up to here": an upload only
  // waits for those (pack_reads for the read buffers, pack_targets for the target buffers; every
  // other reader of tg_off is followed by a host synchronisation before its call returns), so a
  // target upload overlaps with the key-table build of the reads that were just set.
  cudaEvent_t ev_rd_free = nullptr, ev_tg_free = nullptr;
  bool pend_reads = false, pend_targets = false;  // enqueued builds whose counters/timers are not read yet
  cudaEvent_t ev[EV_COUNT] = {};
  std::string err;
  msc_stats st{};

  // reads
  uint64_t n_reads = 0;
  bool have_reads = false;
  DevBuf rd_ascii, rd_offs, rd_words, rd_x, len_flags, validmask, rmeta;
  // key table (common.cuh / build.cuh): 128-byte buckets, partitioned build
  TableGeom tgeo{};
  uint64_t n_slots = 0;  // kBucketSlots * n_buckets
  int lg_bloom = 0;
  BloomGeom geom{};
  uint64_t n_keys = 0, n_groups = 0, n_dup = 0;
  DevBuf tab, recs, dups, part_count, pass_small, pass_cnt, bloom, items;
  int lg_small = 22;          // hashed passing-pair counters of the MaxMatches pre-check (16 MB: L2-resident)
  bool exact_counts = false;  // the pair kernel also keeps exact per-slot counts (after a small counter exceeded MaxMatches)
  // zero-fills already issued by a merged prologue launch (consumed by the stage that owns them)
  struct { bool reads = false, targets = false, scan = false, pairs = false, combine = false; } pro;
  // targets
  uint64_t n_targets = 0, n_bases = 0, n_words_alloc = 0, n_tiles = 0;
  bool have_targets = false;
  bool targets_packed = false;  // set by msc_set_targets_packed: no ASCII copy, rebuild(2) has nothing to redo
  DevBuf tg_ascii, tg_off, tg_words, tg_x, xsum, blk2gene;
  // candidates / pairs
  uint64_t n_cand = 0, n_pairs = 0;
  bool have_cand = false;
  DevBuf cinfo, sizes, pstart, block_first;
  // matches
  uint64_t n_match_pre = 0, n_match = 0;
  bool have_confirm = false, have_combine = false;
  // MSC_STAGE_DEFER: screen + confirm were enqueued without a host synchronisation; they are
  // completed (counters read, buffers checked) by the next call that synchronises
  bool deferred = false;
  DevBuf match_pre, best, rcount, rstart, rfill, match_out, long_list, mid_list;
  // targets sharded over several contexts: MaxMatches is detected at MaxMatches / n_shards and
  // resolved by the host protocol (msc_overflow_keys / msc_divert_groups / msc_replay_diverted)
  int n_shards = 1;
  bool shard_overflow = false;
  // confirm kernel mode 2 (MaxMatches overflow groups diverted to the host)
  struct {
    const uint8_t* slot_over = nullptr;
    uint4* over = nullptr;
    uint64_t over_cap = 0;
  } pair_mode2;
  // device-side read prep (msc_prep_reads): sorted permutation of the raw reads and group starts
  uint64_t prep_kept = 0, prep_unique = 0, prep_bytes = 0;
  bool have_prep = false;
  DevBuf prep_perm, prep_gstart, nm_flag, nm_pos, nm_list;
  struct PrepTmp {  // scratch of msc_prep_reads, kept between calls
    DevBuf d_raw, d_offs, planes, key64, idx_a, idx_b, keep, hist, hoff, head, head_scan, ulen, uoffs;
    void release_all() {
      DevBuf* t[] = {&d_raw, &d_offs, &planes, &key64, &idx_a, &idx_b, &keep, &hist, &hoff, &head, &head_scan, &ulen, &uoffs};
      for (DevBuf* b : t) b->release();
    }
  } prep;
  cudaEvent_t ev_prep0 = nullptr, ev_prep1 = nullptr;
  // misc
  DevBuf counters, tile_sums, scan_state, nmiss;
  unsigned long long scan_arrivals = 0;  // arrivals the scan kernel's grid barrier has seen so far (never reset)
  unsigned long long* h_counters = nullptr;  // pinned mirror
  // MSC_TRACE=1: an event after every launch, per-launch device times printed at each sync
  // stage boundary events other than the scan kernel's pair: MSC_STAGE_EVENTS=0 leaves them out
  // (every record is one more stream operation between two kernels)
  bool stage_events = true;
  bool pdl = true;  // MSC_PDL=0: plain stream-ordered launches
  bool pdl_always = false;  // MSC_PDL=2
  // Programmatic dependent launch hides ~1 us per kernel boundary -- 20 us of a 0.6 ms step on the small workload --
  // but was measured to COST 7.6 ms per step at configs[2] size (key-table build 47.2 -> 39.6 ms without it: the
  // early-launched CTAs of the next kernel take the places of the running kernel's later waves), so it is used only
  // where launch latency matters.
  bool pdl_on() const {
    return pdl && (pdl_always || (n_reads * (uint64_t)std::max(1, win.nwin) <= (1ull << 24) && n_bases <= (1ull << 29)));
  }
  bool trace = false;
  std::vector<std::pair<const char*, cudaEvent_t>> trace_ev;
  size_t trace_used = 0;
  void trace_mark(const char* what) {
    if (trace_used == trace_ev.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      trace_ev.emplace_back(what, e);
    }
    trace_ev[trace_used].first = what;
    cudaEventRecord(trace_ev[trace_used].second, stream);
    trace_used++;
  }
  void trace_dump() {
    for (size_t i = 1; i < trace_used; i++) {
      float ms = 0;
      cudaEventElapsedTime(&ms, trace_ev[i - 1].second, trace_ev[i].second);
      fprintf(stderr, "[msc trace] %8.1f us  %s\n", ms * 1000.f, trace_ev[i].first);
    }
    if (trace_used) fprintf(stderr, "[msc trace] ---- sync\n");
    trace_used = 0;
  }

  int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
  unsigned long long* ctr(int which) const { return counters.as<unsigned long long>() + which; }
  uint64_t cand_cap() const { return sizes.cap >= sizeof(uint32_t) ? sizes.cap / sizeof(uint32_t) - 1 : 0; }
  uint64_t match_cap() const { return match_pre.cap / sizeof(uint4); }
  uint64_t block_cap() const {
    const uint64_t n = block_first.cap / sizeof(uint32_t);
    return n >= 2 ? n - 2 : 0;
  }
};

#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return ctx->fail(MSC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

#define LAUNCH_CHECK()                                       \
  do {                                                       \
    ctx->st.kernel_launches++;                               \
    CK(cudaGetLastError());                                  \
    if (ctx->trace) ctx->trace_mark(__FILE__ ":" MSC_STR(__LINE__)); \
  } while (0)
#define MSC_STR2(x) #x
#define MSC_STR(x) MSC_STR2(x)

#define RC(call)           \
  do {                     \
    int rc__ = (call);     \
    if (rc__) return rc__; \
  } while (0)

namespace {

inline unsigned grid_for(uint64_t n, int block) { return (unsigned)std::max<uint64_t>(1, (n + block - 1) / block); }

int ceil_log2(uint64_t v) {
  int l = 0;
  while ((1ull << l) < v) l++;
  return l;
}

// Exclusive scan of uint32 in[] -> OutT out[] (+ out[n] = total when write_end).  The element
// count is n_host, or the device counter *n_ptr clamped to n_host.  The grand total goes to
// *total (a device counter).  One launch, no host involvement.
template <typename OutT>
int enqueue_exclusive_scan(msc_ctx* ctx, const uint32_t* in, const unsigned long long* n_ptr, uint64_t n_host, OutT* out,
                           bool write_end, unsigned long long* total) {
  // one block per SM: the whole grid is resident, which the kernel's grid barrier relies on
  const unsigned grid = (unsigned)std::min(ctx->sm_count, kScanThreads);
  if (ctx->tile_sums.cap == 0) CK(ctx->tile_sums.reserve((size_t)kScanThreads * sizeof(uint64_t)));
  // COOPERATIVE launch: the runtime starts the grid only when all of its blocks can be resident at
  // the same time (or fails the launch), so the barrier cannot deadlock when other contexts or
  // processes share the GPU (up to MaxConfirmProcs concurrent callers, cmd/muscato/main.go:391-420).
  // The arrival target is advanced only once the launch is known to have been accepted.
  {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kScanThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const unsigned long long target = ctx->scan_arrivals + grid;
    cudaError_t e = cudaLaunchKernelEx(&cfg, scan_resident_kernel<OutT>, in, n_ptr, (uint64_t)n_host, out, write_end ? 1 : 0, total,
                                       ctx->tile_sums.as<uint64_t>(), ctx->scan_state.as<unsigned long long>(), target);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      return ctx->fail(MSC_ERR_CUDA, "cooperative launch of the resident scan failed: %s", cudaGetErrorString(e));
    }
    ctx->scan_arrivals = target;
  }
  LAUNCH_CHECK();
  return MSC_OK;
}

// Fill several buffers in one launch.  Sizes are rounded up to 16 bytes: every DevBuf is
// over-allocated by at least 256 bytes, so the round-up stays inside the allocation.
struct Filler {
  FillJob job{};
  bool overflow = false;  // more than kMaxFillJobs jobs: reported by enqueue_fill as an error
  void add(void* p, size_t bytes, unsigned int value = 0) {
    if (job.n >= kMaxFillJobs) {
      overflow = true;
      return;
    }
    job.ptr[job.n] = p;
    job.bytes[job.n] = (bytes + 15) & ~(size_t)15;
    job.value[job.n] = value;
    job.n++;
  }
};

// The zero-fills every stage needs before it runs.  A fused run issues them all in ONE prologue
// launch (add_*_fills + ctx->pro flags); a stage enqueued on its own issues just its own.
void add_reads_fills(msc_ctx* ctx, Filler& f) {
  f.add(ctx->len_flags.p, (ctx->n_reads + 1) * sizeof(uint32_t));
  f.add(ctx->ctr(C_NKEYS), 4 * sizeof(unsigned long long));  // C_NKEYS, C_NGROUPS, C_NDUP, C_SCRATCH
  f.add(ctx->part_count.p, (size_t)kMaxParts * kPartStride * sizeof(unsigned int));
  f.add(ctx->bloom.p, (1ull << ctx->lg_bloom) * sizeof(uint64_t));  // build_windows_kernel sets the bits
  // (the table's fingerprints are cleared by table_clear_kernel: 64 of every 128 bytes)
}
void add_targets_fills(msc_ctx* ctx, Filler& f) {
  f.add(ctx->tg_x.p, ctx->n_words_alloc * sizeof(uint64_t));
  f.add(ctx->xsum.p, (ctx->n_words_alloc / 32 + 4) * sizeof(uint32_t));
  f.add(ctx->ctr(C_TGX), 2 * sizeof(unsigned long long));
}
void add_scan_fills(msc_ctx* ctx, Filler& f) {
  f.add(ctx->ctr(C_NCAND), 2 * sizeof(unsigned long long));  // C_NCAND, C_BLOOMPASS
  f.add(ctx->ctr(C_NPAIRS), 2 * sizeof(unsigned long long));  // C_NPAIRS (rewritten by every expansion), C_NDUMMY
}
void add_pairs_fills(msc_ctx* ctx, Filler& f) {
  f.add(ctx->ctr(C_NMATCH), 4 * sizeof(unsigned long long));  // C_NMATCH, C_NPASS, C_NOVER, C_NOUT
  f.add(ctx->best.p, (ctx->n_reads + 1) * sizeof(uint32_t), MSC_NO_MATCH);
  f.add(ctx->pass_small.p, (1ull << ctx->lg_small) * sizeof(uint32_t));
  if (ctx->exact_counts) f.add(ctx->pass_cnt.p, ctx->n_slots * sizeof(uint32_t));
}
void add_combine_fills(msc_ctx* ctx, Filler& f) {
  f.add(ctx->rcount.p, (ctx->n_reads + 1) * sizeof(uint32_t));
  f.add(ctx->ctr(C_NLONG), 2 * sizeof(unsigned long long));
}

int enqueue_fill(msc_ctx* ctx, const Filler& f) {
  if (f.overflow) return ctx->fail(MSC_ERR_STATE, "internal: more than %d fill jobs in one prologue", kMaxFillJobs);
  if (f.job.n == 0) return MSC_OK;
  launch_k(ctx->pdl_on(), fill_buffers_kernel, (unsigned)ctx->sm_count * 8, 256, 0, ctx->stream, f.job);
  LAUNCH_CHECK();
  return MSC_OK;
}

void account_build_reads(msc_ctx* ctx);
void account_pack_targets(msc_ctx* ctx);

// The one host synchronisation of a call: brings the counter block over and books the
// counters / timers of every build that was enqueued since the last one.
int sync_counters(msc_ctx* ctx) {
  CK(cudaMemcpyAsync(ctx->h_counters, ctx->counters.p, C_COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                     ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->st.d2h_bytes += C_COUNT * sizeof(unsigned long long);
  if (ctx->pend_reads) account_build_reads(ctx);
  if (ctx->pend_targets) account_pack_targets(ctx);
  ctx->pend_reads = ctx->pend_targets = false;
  if (ctx->trace) ctx->trace_dump();
  return MSC_OK;
}

// Input upload on the copy stream.  The copy waits for the kernels that may still read its
// destination (ev_rd_free / ev_tg_free); the compute stream
// waits for the copy; the HOST only waits for the copy, because the caller's buffers are
// borrowed for the duration of the call -- the kernels that consume the data keep running while
// the caller prepares (or uploads) its next input.
int begin_upload(msc_ctx* ctx, cudaEvent_t dest_free) {
  CK(cudaStreamWaitEvent(ctx->copy_stream, dest_free, 0));
  return MSC_OK;
}
int end_upload(msc_ctx* ctx) {
  CK(cudaEventRecord(ctx->ev_copy, ctx->copy_stream));
  CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy, 0));
  return MSC_OK;
}
int wait_upload(msc_ctx* ctx) {
  CK(cudaEventSynchronize(ctx->ev_copy));
  return MSC_OK;
}
int wait_upload_all(msc_ctx* ctx) {  // error paths: nothing of the caller's buffers may still be in flight
  CK(cudaStreamSynchronize(ctx->copy_stream));
  return MSC_OK;
}

float elapsed(msc_ctx* ctx, int a, int b) {
  float ms = 0;
  if (cudaEventElapsedTime(&ms, ctx->ev[a], ctx->ev[b]) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0.f;
  }
  return ms;
}

// ---- enqueue: reads pack + key table build (from the resident ASCII copy) ------------------
// Three parts, so that msc_set_reads can run the per-read work chunk by chunk behind a chunked upload:
//   begin   zero-fills (+ the table clear when uploading: it then runs under the first chunk's copy)
//   chunk   2-bit pack + window pass (validity, read records, front bits, keys per partition) of reads [r0, r1)
//   end     partition offsets -> scatter -> insert -> CSR of the further group members
static BuildArgs build_args(msc_ctx* ctx) {
  BuildArgs a{};
  a.rd_words = ctx->rd_words.as<uint64_t>();
  a.rd_words_rw = ctx->rd_words.as<uint64_t>();
  a.rd_x = ctx->rd_x.as<uint64_t>();
  a.len_flags = ctx->len_flags.as<uint32_t>();
  a.n_reads = ctx->n_reads;
  a.w_begin = 0;
  a.w_end = ctx->n_reads;
  a.nmiss = ctx->nmiss.as<int32_t>();
  a.validmask = ctx->validmask.as<uint32_t>();
  a.rmeta = ctx->rmeta.as<uint2>();
  a.n_keys = ctx->ctr(C_NKEYS);
  a.tab = ctx->tab.as<uint8_t>();
  a.tg = ctx->tgeo;
  a.part_count = ctx->part_count.as<unsigned int>();
  a.recs = ctx->recs.as<uint4>();
  a.dups = ctx->dups.as<uint4>();
  a.n_dup = ctx->ctr(C_NDUP);
  a.n_alloc = ctx->ctr(C_SCRATCH);
  a.insert_cursor = ctx->ctr(C_NGROUPS);  // (the host derives the group count from keys - further members)
  a.items = ctx->items.as<uint2>();
  a.bloom = ctx->bloom.as<unsigned long long>();
  a.geom = ctx->geom;
  return a;
}

static int enqueue_table_clear(msc_ctx* ctx) {
  launch_k(ctx->pdl_on(), table_clear_kernel, (unsigned)std::min<uint64_t>(grid_for(ctx->tgeo.n_buckets * 4, 256), (uint64_t)ctx->sm_count * 32),
           256, 0, ctx->stream, ctx->tab.as<uint8_t>(), ctx->tgeo.n_buckets);
  LAUNCH_CHECK();
  return MSC_OK;
}

int enqueue_build_begin(msc_ctx* ctx, bool clear_now) {
  if (ctx->stage_events) CK(cudaEventRecord(ctx->ev[EV_PACKR0], ctx->stream));
  if (ctx->trace) ctx->trace_mark("start build_reads");
  if (!ctx->pro.reads) {
    Filler f;
    add_reads_fills(ctx, f);
    RC(enqueue_fill(ctx, f));
  }
  ctx->pro.reads = false;
  if (clear_now && ctx->n_reads) RC(enqueue_table_clear(ctx));
  return MSC_OK;
}

int enqueue_pack_reads_range(msc_ctx* ctx, uint64_t r0, uint64_t r1) {
  if (r1 <= r0) return MSC_OK;
  const int S = ctx->win.S;
  const int rpb = std::max(1, 256 / S);  // whole reads per block
  const size_t smem = (size_t)rpb * (size_t)ctx->win.MRL + 64;
  launch_k(ctx->pdl_on(), pack_reads_kernel, grid_for(r1 - r0, rpb), 256, smem, ctx->stream,
      ctx->rd_ascii.as<uint8_t>(), ctx->rd_offs.as<uint64_t>() + r0, r1 - r0, S, rpb, ctx->rd_words.as<uint64_t>() + r0 * (uint64_t)S,
      ctx->rd_x.as<uint64_t>() + r0 * (uint64_t)S, ctx->len_flags.as<uint32_t>() + r0);
  LAUNCH_CHECK();
  return MSC_OK;
}

int enqueue_windows_range(msc_ctx* ctx, uint64_t r0, uint64_t r1) {
  if (r1 <= r0) return MSC_OK;
  BuildArgs a = build_args(ctx);
  a.w_begin = r0;
  a.w_end = r1;
  const unsigned g8 = (unsigned)ctx->sm_count * 8;
  launch_k(ctx->pdl_on(), build_windows_kernel, (unsigned)std::min<uint64_t>(grid_for(r1 - r0, 256), g8), 256, 0, ctx->stream, ctx->win, a);
  LAUNCH_CHECK();
  return MSC_OK;
}

int enqueue_build_end(msc_ctx* ctx, bool clear_now) {
  const uint64_t U = ctx->n_reads;
  if (U) {
    const BuildArgs a = build_args(ctx);
    const unsigned g8 = (unsigned)ctx->sm_count * 8;
    const uint64_t n_items = U * (uint64_t)ctx->win.nwin;
    if (clear_now) RC(enqueue_table_clear(ctx));
    launch_k(ctx->pdl_on(), build_offsets_kernel, 1, kMaxParts, 0, ctx->stream, ctx->part_count.as<unsigned int>(), (int)ctx->tgeo.n_parts);
    LAUNCH_CHECK();
    if (!ctx->scatter_attr_set) {
      CK(cudaFuncSetAttribute((const void*)build_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterSmem)));
      ctx->scatter_attr_set = true;
    }
    launch_k(ctx->pdl_on(), build_scatter_kernel, (unsigned)std::min<uint64_t>(grid_for(n_items, kStageKeys), (uint64_t)ctx->sm_count * 2),
             kScatterThreads, sizeof(ScatterSmem), ctx->stream, ctx->win, a);
    LAUNCH_CHECK();
    launch_k(ctx->pdl_on(), build_insert_kernel, (unsigned)std::min<uint64_t>(grid_for(n_items, 256), g8), 256, 0, ctx->stream, a);
    LAUNCH_CHECK();
    launch_k(ctx->pdl_on(), build_dup_count_kernel, g8, 256, 0, ctx->stream, a);
    LAUNCH_CHECK();
    launch_k(ctx->pdl_on(), build_dup_alloc_kernel, g8, 256, 0, ctx->stream, a);
    LAUNCH_CHECK();
    launch_k(ctx->pdl_on(), build_dup_fill_kernel, g8, 256, 0, ctx->stream, a);
    LAUNCH_CHECK();
  }
  if (ctx->stage_events) CK(cudaEventRecord(ctx->ev[EV_BUILD1], ctx->stream));
  ctx->have_cand = ctx->have_confirm = ctx->have_combine = false;
  ctx->pend_reads = true;
  return MSC_OK;
}

// The whole build from the resident ASCII copy (msc_rebuild, device-side inputs).
int enqueue_build_reads(msc_ctx* ctx) {
  RC(enqueue_build_begin(ctx, false));
  RC(enqueue_pack_reads_range(ctx, 0, ctx->n_reads));
  CK(cudaEventRecord(ctx->ev_rd_free, ctx->stream));
  if (ctx->stage_events) CK(cudaEventRecord(ctx->ev[EV_PACKR1], ctx->stream));
  RC(enqueue_windows_range(ctx, 0, ctx->n_reads));
  return enqueue_build_end(ctx, true);
}

void account_build_reads(msc_ctx* ctx) {
  ctx->n_keys = ctx->h_counters[C_NKEYS];
  ctx->n_dup = ctx->h_counters[C_NDUP];
  ctx->n_groups = ctx->n_keys - ctx->n_dup;
  ctx->st.n_reads = ctx->n_reads;
  ctx->st.n_keys = ctx->n_keys;
  ctx->st.n_key_groups = ctx->n_groups;
  ctx->st.table_slots = ctx->n_slots;
  ctx->st.bloom_bytes = (1ull << ctx->lg_bloom) * sizeof(uint64_t);
  if (ctx->stage_events) {
    ctx->st.ms_pack_reads += elapsed(ctx, EV_PACKR0, EV_PACKR1);
    ctx->st.ms_build += elapsed(ctx, EV_PACKR1, EV_BUILD1);
  }
}

int enqueue_pack_targets(msc_ctx* ctx) {
  if (ctx->stage_events) CK(cudaEventRecord(ctx->ev[EV_PACKT0], ctx->stream));
  if (!ctx->pro.targets) {
    Filler f;
    add_targets_fills(ctx, f);
    RC(enqueue_fill(ctx, f));
  }
  ctx->pro.targets = false;
  launch_k(ctx->pdl_on(), pack_targets_kernel, grid_for(ctx->n_words_alloc, 256), 256, 0, ctx->stream, 
      ctx->tg_ascii.as<uint8_t>(), ctx->n_bases, ctx->tg_words.as<uint64_t>(), ctx->n_words_alloc,
      ctx->tg_x.as<uint64_t>(), ctx->xsum.as<uint32_t>(), ctx->ctr(C_TGX));
  LAUNCH_CHECK();
  CK(cudaEventRecord(ctx->ev_tg_free, ctx->stream));
  if (ctx->stage_events) CK(cudaEventRecord(ctx->ev[EV_PACKT1], ctx->stream));
  ctx->have_cand = ctx->have_confirm = ctx->have_combine = false;
  ctx->pend_targets = true;
  return MSC_OK;
}

void account_pack_targets(msc_ctx* ctx) {
  if (ctx->stage_events) ctx->st.ms_pack_targets += elapsed(ctx, EV_PACKT0, EV_PACKT1);
  ctx->st.n_targets = ctx->n_targets;
  ctx->st.target_bases = ctx->n_bases;
}

// The candidate list is two parallel arrays: (slot, position) and the slot's record.
int reserve_cand(msc_ctx* ctx, uint64_t n) {
  // per candidate: the pair kernel's 32-byte record (position, slot, group record ...) and the key group's size
  CK(ctx->sizes.reserve((n + 1) * sizeof(uint32_t)));
  const uint64_t ccap = ctx->sizes.cap / sizeof(uint32_t) - 1;
  CK(ctx->cinfo.reserve((ccap + 1) * 2 * sizeof(uint4)));
  return MSC_OK;
}

// ---- enqueue: scan ---------------------------------------------------------------------------
template <int KW>
void (*pick_scan_wn(int wn))(const ScanArgs) {
  switch (wn) {
    case 1: return scan_targets_kernel<KW, 1>;
    case 2: return scan_targets_kernel<KW, 2>;
    case 3: return scan_targets_kernel<KW, 3>;
    case 4: return scan_targets_kernel<KW, 4>;
    case 5: return scan_targets_kernel<KW, 5>;
    case 6: return scan_targets_kernel<KW, 6>;
    case 7: return scan_targets_kernel<KW, 7>;
    default: return scan_targets_kernel<KW, 8>;
  }
}
// W <= 16: 32-bit keys; W <= 32: one 64-bit key word; wider: two key words
void (*pick_scan_kernel(int W, int wn, bool direct = false))(const ScanArgs) {
  if (direct) return scan_direct_kernel;  // exact front (W <= 15), scan_direct.cuh
  return W <= 16 ? pick_scan_wn<0>(wn) : W <= 32 ? pick_scan_wn<1>(wn) : pick_scan_wn<2>(wn);
}

int enqueue_scan(msc_ctx* ctx) {
  if (ctx->sizes.cap == 0) {
    // (exact front: every warp of every launch may leave most of one kSlotBlock of reserved slots empty)
    const uint64_t slack = ctx->geom.direct ? (uint64_t)ctx->sm_count * MSC_DIRECT_CTAS * kScanWarps * kSlotBlock << ctx->geom.lg_pass : 0;
    RC(reserve_cand(ctx, std::max<uint64_t>(1u << 20, ctx->n_bases / 16) + slack));
  }
  void (*scan_fn)(const ScanArgs) = pick_scan_kernel(ctx->win.W, ctx->geom.wn, ctx->geom.direct != 0);
  const size_t scan_smem = ctx->geom.direct ? sizeof(ScanDirectSmem) : sizeof(ScanSmem);
  if (ctx->scan_grid == 0 || ctx->scan_fn_sized != (const void*)scan_fn) {
    int blocks_per_sm = 0;
    CK(cudaFuncSetAttribute((const void*)scan_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, scan_fn, kScanBlock, scan_smem));
    ctx->scan_grid = ctx->sm_count * std::max(1, blocks_per_sm);
    ctx->scan_fn_sized = (const void*)scan_fn;
  }
  if (!ctx->pro.scan) {
    Filler f;
    add_scan_fills(ctx, f);
    RC(enqueue_fill(ctx, f));
  }
  ctx->pro.scan = false;
  CK(cudaEventRecord(ctx->ev[EV_SCAN0], ctx->stream));
  if (ctx->n_tiles && ctx->n_reads) {
    ScanArgs a{};
    a.tg_words = ctx->tg_words.as<uint64_t>();
    a.tg_x = ctx->tg_x.as<uint64_t>();
    a.xsum = ctx->xsum.as<uint32_t>();
    a.targets_have_x = ctx->ctr(C_TGX);
    a.n_bases = ctx->n_bases;
    a.n_tiles = ctx->n_tiles;
    a.bloom = ctx->bloom.as<uint2>();
    a.geom = ctx->geom;
    for (int j = 0; j < 8; j++) a.mul[j] = j < ctx->geom.wn ? 1u << (32 - 2 * ctx->geom.m - 2 * j) : 0u;
    a.tab = ctx->tab.as<uint8_t>();
    a.n_buckets = ctx->tgeo.n_buckets;
    a.cinfo = ctx->cinfo.as<uint4>();
    a.sizes = ctx->sizes.as<uint32_t>();
    a.tg_off = ctx->tg_off.as<uint32_t>();
    a.blk2gene = ctx->blk2gene.as<uint32_t>();
    a.n_targets = ctx->n_targets;
    a.cand_cap = ctx->cand_cap();
    a.n_cand = ctx->ctr(C_NCAND);
    a.n_bloom_pass = ctx->ctr(C_BLOOMPASS);
    a.n_dummy = ctx->ctr(C_NDUMMY);
    a.W = ctx->win.W;
    a.alu_masks = ctx->lg_bloom > 23 ? 1 : 0;
    if (const char* e = getenv("MSC_SCAN_ALU_MASKS")) a.alu_masks = atoi(e) != 0;
    // Bloom front: measured slower at S2 (72 vs 60 ms: the memory system is saturated by the filter's own misses, more
    // requests in flight only add queueing); exact front: the survivor's line is requested when it is queued
    a.prefetch = 0;
    if (const char* e = getenv("MSC_SCAN_PREFETCH")) a.prefetch = atoi(e);
    // table beyond the L2: its lines and the candidate records are touched once (evict-first), which leaves the L2 to
    // the front's current slice
    a.stream_tab = ctx->tgeo.n_buckets * (uint64_t)kBucketBytes > (64ull << 20) ? 1 : 0;
    if (const char* e = getenv("MSC_SCAN_STREAM_TAB")) a.stream_tab = atoi(e) != 0;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(ctx->n_tiles, (uint64_t)ctx->scan_grid));
    // exact front: one launch per slice of the bitmap (every launch walks the whole database and tests the positions
    // whose bit lies in its slice); the Bloom front is one launch
    const int n_pass = ctx->geom.direct ? 1 << ctx->geom.lg_pass : 1;
    for (int q = 0; q < n_pass; q++) {
      a.pass = q;
      launch_k(ctx->pdl_on(), scan_fn, grid, kScanBlock, scan_smem, ctx->stream, a);
      LAUNCH_CHECK();
    }
  }
  CK(cudaEventRecord(ctx->ev[EV_SCAN1], ctx->stream));
  return MSC_OK;
}

// ---- enqueue: expansion + pair kernel -------------------------------------------------------
int enqueue_pairs(msc_ctx* ctx, int mode, DevBuf& outbuf) {
  const uint64_t ccap = ctx->cand_cap();
  CK(ctx->pstart.reserve((ccap + 2) * sizeof(uint64_t)));
  if (ctx->block_first.cap == 0) CK(ctx->block_first.reserve(((size_t)(1u << 19) + 2) * sizeof(uint32_t)));
  if (outbuf.cap == 0) CK(outbuf.reserve((size_t)(1u << 20) * sizeof(uint4)));
  const unsigned pgrid = (unsigned)ctx->sm_count * 8;
  RC(enqueue_exclusive_scan<uint64_t>(ctx, ctx->sizes.as<uint32_t>(), ctx->ctr(C_NCAND), ccap, ctx->pstart.as<uint64_t>(),
                                      true, ctx->ctr(C_NPAIRS)));
  launch_k(ctx->pdl_on(), pair_block_starts_kernel, pgrid, 256, 0, ctx->stream, ctx->pstart.as<uint64_t>(), ctx->ctr(C_NCAND), ccap,
                                                           ctx->ctr(C_NPAIRS), ctx->block_cap(),
                                                           ctx->block_first.as<uint32_t>());
  LAUNCH_CHECK();
  if (ctx->stage_events) CK(cudaEventRecord(ctx->ev[EV_EXPAND1], ctx->stream));

  if (!ctx->pro.pairs) {
    Filler f;
    add_pairs_fills(ctx, f);
    RC(enqueue_fill(ctx, f));
  }
  ctx->pro.pairs = false;
  ConfirmArgs a{};
  a.cinfo = ctx->cinfo.as<uint4>();
  a.block_first = ctx->block_first.as<uint32_t>();
  a.pstart = ctx->pstart.as<uint64_t>();
  a.n_pairs_ptr = ctx->ctr(C_NPAIRS);
  a.block_cap = ctx->block_cap();
  a.items = ctx->items.as<uint2>();
  a.validmask = ctx->validmask.as<uint32_t>();
  a.pass_small = ctx->pass_small.as<uint32_t>();
  a.lg_small = ctx->lg_small;
  a.pass_cnt = ctx->exact_counts ? ctx->pass_cnt.as<uint32_t>() : nullptr;
  a.rd_words = ctx->rd_words.as<uint64_t>();
  a.rd_x = ctx->rd_x.as<uint64_t>();
  a.tg_words = ctx->tg_words.as<uint64_t>();
  a.tg_x = ctx->tg_x.as<uint64_t>();
  a.xsum = ctx->xsum.as<uint32_t>();
  a.tg_off = ctx->tg_off.as<uint32_t>();
  a.n_targets = ctx->n_targets;
  a.nwin_magic = ctx->win.nwin == 1 ? 0ull : (~0ull / (uint64_t)ctx->win.nwin) + 1ull;
  a.targets_have_x = ctx->ctr(C_TGX);
  a.matches = outbuf.as<uint4>();
  a.match_cap = outbuf.cap / sizeof(uint4);
  a.n_match = ctx->ctr(C_NMATCH);
  a.n_pass = ctx->ctr(C_NPASS);
  a.best = ctx->best.as<uint32_t>();
  a.mode = mode;
  if (mode == 2) {
    a.slot_over = ctx->pair_mode2.slot_over;
    a.tab = ctx->tab.as<uint8_t>();
    a.n_buckets = ctx->tgeo.n_buckets;
    a.over = ctx->pair_mode2.over;
    a.over_cap = ctx->pair_mode2.over_cap;
    a.n_over_inst = ctx->ctr(C_NLONG);
  }
  if (ctx->confirm_grid == 0) {  // persistent grid: exactly the resident CTAs
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, confirm_pairs_kernel<0>, 256, 0));
    ctx->confirm_grid = ctx->sm_count * std::max(1, per_sm);
  }
  const unsigned cgrid = (unsigned)ctx->confirm_grid;
  if (mode == 0) launch_k(ctx->pdl_on(), confirm_pairs_kernel<0>, cgrid, 256, 0, ctx->stream, ctx->win, a);
  else if (mode == 1) launch_k(ctx->pdl_on(), confirm_pairs_kernel<1>, cgrid, 256, 0, ctx->stream, ctx->win, a);
  else launch_k(ctx->pdl_on(), confirm_pairs_kernel<2>, cgrid, 256, 0, ctx->stream, ctx->win, a);
  LAUNCH_CHECK();
  if (mode == 0) {
    // MaxMatches pre-check (cmd/muscato_confirm/main.go:233-242, :424-448): truncation can only
    // happen in a key group with more than MaxMatches passing pairs.
    const unsigned long long thr = ctx->n_shards > 1 ? (unsigned long long)ctx->cfg.max_matches / (unsigned long long)ctx->n_shards
                                                     : (unsigned long long)ctx->cfg.max_matches;
    // exact per-slot counts once they are kept, else the small hashed counters (upper bounds)
    launch_k(ctx->pdl_on(), overflow_count_kernel, (unsigned)ctx->sm_count * 8, 256, 0, ctx->stream,
        ctx->exact_counts ? ctx->pass_cnt.as<uint32_t>() : ctx->pass_small.as<uint32_t>(),
        ctx->exact_counts ? ctx->n_slots : (1ull << ctx->lg_small), thr, ctx->ctr(C_NPASS), ctx->ctr(C_NOVER),
        ctx->n_shards > 1 ? ctx->best.as<uint32_t>() + ctx->n_reads : (uint32_t*)nullptr);
    LAUNCH_CHECK();
  }
  if (ctx->stage_events) CK(cudaEventRecord(ctx->ev[EV_CONFIRM1], ctx->stream));
  return MSC_OK;
}

int enqueue_combine(msc_ctx* ctx) {
  const uint64_t U = ctx->n_reads, mcap = ctx->match_cap();
  CK(ctx->rcount.reserve((U + 1) * sizeof(uint32_t)));
  CK(ctx->rstart.reserve((U + 2) * sizeof(uint32_t)));
  CK(ctx->rfill.reserve((U + 1) * sizeof(uint32_t)));
  CK(ctx->match_out.reserve((mcap + 1) * sizeof(uint4)));
  CK(ctx->long_list.reserve((U + 1) * sizeof(uint32_t)));
  CK(ctx->mid_list.reserve((U + 1) * sizeof(uint32_t)));
  if (ctx->stage_events) CK(cudaEventRecord(ctx->ev[EV_COMB0], ctx->stream));
  if (!ctx->pro.combine) {
    Filler f;
    add_combine_fills(ctx, f);
    RC(enqueue_fill(ctx, f));
  }
  ctx->pro.combine = false;
  const unsigned g = (unsigned)ctx->sm_count * 8;
  if (ctx->n_shards > 1) {  // best[n_reads] after the MIN all-reduce: did ANY shard see a MaxMatches candidate group?
    launch_k(ctx->pdl_on(), shard_flag_kernel, 1, 32, 0, ctx->stream, ctx->best.as<uint32_t>() + ctx->n_reads, ctx->ctr(C_SHARDOVER));
    LAUNCH_CHECK();
  }
  launch_k(ctx->pdl_on(), combine_count_kernel, g, 256, 0, ctx->stream, ctx->match_pre.as<uint4>(), ctx->ctr(C_NMATCH), mcap,
                                                   ctx->best.as<uint32_t>(), (uint32_t)ctx->cfg.mmtol,
                                                   ctx->rcount.as<uint32_t>());
  LAUNCH_CHECK();
  RC(enqueue_exclusive_scan<uint32_t>(ctx, ctx->rcount.as<uint32_t>(), nullptr, U, ctx->rstart.as<uint32_t>(), true,
                                      ctx->ctr(C_NOUT)));
  // the per-read output cursors start at the reads' first slots
  CK(cudaMemcpyAsync(ctx->rfill.p, ctx->rstart.p, (U + 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  launch_k(ctx->pdl_on(), combine_scatter_kernel, g, 256, 0, ctx->stream, ctx->match_pre.as<uint4>(), ctx->ctr(C_NMATCH), mcap,
                                                     ctx->best.as<uint32_t>(), (uint32_t)ctx->cfg.mmtol,
                                                     ctx->rfill.as<uint32_t>(), ctx->match_out.as<uint4>());
  LAUNCH_CHECK();
  if (U) {
    // deterministic (gene, pos) order inside every read group; match_pre is dead and serves as scratch
    launch_k(ctx->pdl_on(), segment_sort_short_kernel, grid_for(U, 256), 256, 0, ctx->stream, 
        ctx->match_out.as<uint4>(), ctx->rstart.as<uint32_t>(), U, ctx->long_list.as<uint32_t>(), ctx->ctr(C_NLONG),
        ctx->mid_list.as<uint32_t>(), ctx->ctr(C_PAD1));
    LAUNCH_CHECK();
    launch_k(ctx->pdl_on(), segment_rank_sort_kernel, g, kRankThreads, 0, ctx->stream, ctx->match_out.as<uint4>(), ctx->match_pre.as<uint4>(),
                                                         ctx->rstart.as<uint32_t>(), ctx->long_list.as<uint32_t>(),
                                                         ctx->ctr(C_NLONG), ctx->mid_list.as<uint32_t>(), ctx->ctr(C_PAD1));
    LAUNCH_CHECK();
  }
  if (ctx->stage_events) CK(cudaEventRecord(ctx->ev[EV_COMB1], ctx->stream));
  return MSC_OK;
}

// ---- completion: read the counter block, grow buffers that were too small --------------------
// finish_* return MSC_OK, an error, or NEED_RETRY when a buffer was grown and the caller must
// enqueue again.
enum { NEED_RETRY = -1, NEED_RECOMBINE = -2 };

int finish_scan(msc_ctx* ctx) {
  const float ms = elapsed(ctx, EV_SCAN0, EV_SCAN1);
  ctx->st.ms_scan += ms;
  ctx->st.ms_scan_kernel = ms;
  ctx->n_cand = ctx->h_counters[C_NCAND];
  if (ctx->n_cand > ctx->cand_cap()) {
    RC(reserve_cand(ctx, ctx->n_cand));
    ctx->have_cand = false;
    return NEED_RETRY;
  }
  ctx->st.n_candidates = ctx->n_cand - std::min<uint64_t>(ctx->n_cand, ctx->h_counters[C_NDUMMY]);  // without the empty slots of the exact front
  ctx->st.positions_probed = ctx->n_bases;
  ctx->st.reserved_f[0] = (float)ctx->h_counters[C_BLOOMPASS];
  ctx->st.reserved_f[2] = ctx->geom.direct ? (float)(1 << ctx->geom.lg_pass) : 0.f;
  ctx->have_cand = true;
  ctx->have_confirm = ctx->have_combine = false;
  return MSC_OK;
}

int finish_pairs(msc_ctx* ctx, DevBuf& outbuf, uint64_t* n_out) {
  ctx->n_pairs = ctx->h_counters[C_NPAIRS];
  const uint64_t n_blocks = (ctx->n_pairs + kPairBlock - 1) / kPairBlock;
  bool retry = false;
  if (n_blocks > ctx->block_cap()) {
    CK(ctx->block_first.reserve((n_blocks + 2) * sizeof(uint32_t)));
    retry = true;
  }
  const uint64_t n = ctx->h_counters[C_NMATCH];
  if (n > outbuf.cap / sizeof(uint4)) {
    CK(outbuf.reserve(n * sizeof(uint4)));
    retry = true;
  }
  if (retry) return NEED_RETRY;
  *n_out = n;
  return MSC_OK;
}

int mark_expand_start(msc_ctx* ctx);

#include "maxmatches.inc"

int finish_confirm(msc_ctx* ctx) {
  RC(finish_pairs(ctx, ctx->match_pre, &ctx->n_match_pre));
  if (ctx->stage_events) {
    ctx->st.ms_expand += elapsed(ctx, EV_SCAN1, EV_EXPAND1);
    ctx->st.ms_confirm += elapsed(ctx, EV_EXPAND1, EV_CONFIRM1);
  }
  ctx->st.n_pairs = ctx->n_pairs;
  ctx->st.n_pass = ctx->h_counters[C_NPASS];
  ctx->st.n_matches_pre = ctx->n_match_pre;
  ctx->st.n_overflow_groups = ctx->h_counters[C_NOVER];
  if (ctx->st.n_overflow_groups && !ctx->exact_counts) {
    // a small counter exceeds the limit: some key group MAY hold more than MaxMatches (/ n_shards) passing
    // pairs.  Keep exact per-slot counts from now on and run the pair kernel again.
    CK(ctx->pass_cnt.reserve(ctx->n_slots * sizeof(uint32_t)));
    ctx->exact_counts = true;
    ctx->st.n_overflow_groups = 0;
    return NEED_RETRY;
  }
  if (ctx->st.n_overflow_groups && ctx->n_shards <= 1) {
    // Some key group holds more passing pairs than MaxMatches: replay the reference's
    // order-dependent truncation for those groups (rare path, host assisted).
    const float ms_e = ctx->st.ms_expand, ms_c = ctx->st.ms_confirm;
    RC(resolve_maxmatches_overflow(ctx));
    ctx->st.ms_expand = ms_e;
    ctx->st.ms_confirm = ms_c;
    ctx->have_confirm = true;
    ctx->have_combine = false;
    return NEED_RECOMBINE;  // a combine enqueued together with this confirm saw the untruncated set
  }
  ctx->have_confirm = true;
  ctx->have_combine = false;
  return MSC_OK;
}

int finish_combine(msc_ctx* ctx) {
  ctx->n_match = ctx->h_counters[C_NOUT];
  ctx->st.n_matches = ctx->n_match;
  if (ctx->stage_events) ctx->st.ms_combine += elapsed(ctx, EV_COMB0, EV_COMB1);
  ctx->have_combine = true;
  ctx->shard_overflow = ctx->n_shards > 1 && ctx->h_counters[C_SHARDOVER] != 0;
  return MSC_OK;
}

// The expand stage is timed from EV_SCAN1; when it is enqueued on its own (msc_confirm after
// msc_screen) that boundary is re-recorded.
int mark_expand_start(msc_ctx* ctx) {
  CK(cudaEventRecord(ctx->ev[EV_SCAN1], ctx->stream));
  return MSC_OK;
}

// [rebuild +] screen [+ confirm + combine] with a single synchronisation; repeated from the
// scan when a buffer had to grow.
int run_pipeline(msc_ctx* ctx, int rebuild_what, bool do_scan, bool do_confirm, bool do_combine, bool defer = false) {
  if (ctx->deferred) {
    // a deferred screen + confirm is in flight: the only continuation is combine (+ its completion)
    if (!do_combine || do_scan || do_confirm || rebuild_what || defer) {
      ctx->deferred = false;  // drop the pending run: the context goes back to "inputs set"
      RC(sync_counters(ctx));
      ctx->have_cand = ctx->have_confirm = ctx->have_combine = false;
      return ctx->fail(MSC_ERR_STATE, "a deferred screen/confirm was pending: only the combine stage may follow it (dropped)");
    }
    ctx->deferred = false;
    RC(enqueue_combine(ctx));
    RC(sync_counters(ctx));
    int rc = finish_scan(ctx);
    if (rc == MSC_OK) rc = finish_confirm(ctx);
    if (rc == MSC_OK) rc = finish_combine(ctx);
    if (rc == NEED_RETRY || rc == NEED_RECOMBINE) {
      // a buffer was grown, or MaxMatches truncation had to rewrite the matches: what was exchanged
      // between the stages is stale -- the caller repeats the sequence (without MSC_STAGE_DEFER when
      // groups overflow MaxMatches)
      ctx->have_cand = ctx->have_confirm = ctx->have_combine = false;
      return ctx->fail(MSC_ERR_AGAIN, "deferred run must be repeated (%s)",
                       rc == NEED_RETRY ? "an output buffer was grown" : "MaxMatches truncation applies: run without MSC_STAGE_DEFER");
    }
    return rc;
  }
  // Fused run: all zero-fills of the stages below in one prologue launch.
  if ((int)((rebuild_what & 1) != 0) + (int)((rebuild_what & 2) != 0) + (int)do_scan + (int)do_confirm + (int)do_combine > 1) {
    Filler f;
    if (rebuild_what & 1) { add_reads_fills(ctx, f); ctx->pro.reads = true; }
    if (rebuild_what & 2) { add_targets_fills(ctx, f); ctx->pro.targets = true; }
    if (do_scan) { add_scan_fills(ctx, f); ctx->pro.scan = true; }
    if (do_confirm) { add_pairs_fills(ctx, f); ctx->pro.pairs = true; }
    if (do_combine) {
      const uint64_t U = ctx->n_reads;
      CK(ctx->rcount.reserve((U + 1) * sizeof(uint32_t)));
      CK(ctx->rfill.reserve((U + 1) * sizeof(uint32_t)));
      add_combine_fills(ctx, f);
      ctx->pro.combine = true;
    }
    RC(enqueue_fill(ctx, f));
  }
  for (int attempt = 0; attempt < 5; attempt++) {
    if (attempt == 0) {
      if (rebuild_what & 1) RC(enqueue_build_reads(ctx));
      if (rebuild_what & 2) RC(enqueue_pack_targets(ctx));
    }
    if (do_scan) RC(enqueue_scan(ctx));
    if (do_confirm) {
      if (!do_scan) RC(mark_expand_start(ctx));
      RC(enqueue_pairs(ctx, 0, ctx->match_pre));
    }
    if (do_combine) RC(enqueue_combine(ctx));
    if (defer) {  // no host synchronisation: completed by the combine call that follows
      ctx->deferred = true;
      return MSC_OK;
    }
    RC(sync_counters(ctx));
    int rc = MSC_OK;
    if (do_scan) rc = finish_scan(ctx);
    if (rc == MSC_OK && do_confirm) rc = finish_confirm(ctx);
    if (rc == NEED_RECOMBINE) {
      rc = MSC_OK;
      if (do_combine) {
        RC(enqueue_combine(ctx));
        RC(sync_counters(ctx));
      }
    }
    if (rc == MSC_OK && do_combine) rc = finish_combine(ctx);
    if (rc != NEED_RETRY) return rc;
    do_scan = !ctx->have_cand;  // scan again only if the candidate buffer was grown
    if (do_scan || do_confirm) {  // the prologue's zero-fills were consumed: the repeated stages issue their own
      ctx->pro.scan = ctx->pro.pairs = false;
    }
  }
  return ctx->fail(MSC_ERR_NOMEM, "output buffers kept overflowing");
}

}  // namespace

// ===========================================================================================
extern "C" {

const char* msc_version(void) { return "muscato_b200 0.2 (sm_100a)"; }

uint64_t msc_struct_size(int which) {
  switch (which) {
    case 0: return sizeof(msc_config);
    case 1: return sizeof(msc_match);
    case 2: return sizeof(msc_stats);
    case 3: return sizeof(msc_key_rec);
    case 4: return sizeof(msc_cand_rec);
    default: return 0;
  }
}

msc_ctx* msc_create(const msc_config* config, char* errbuf, uint64_t errlen) {
  auto fail = [&](const char* msg) -> msc_ctx* {
    if (errbuf && errlen) snprintf(errbuf, (size_t)errlen, "%s", msg);
    return nullptr;
  };
  if (!config) return fail("config is NULL");
  const msc_config& c = *config;
  // checkArgs (cmd/muscato/main.go:851-858, :871-874): Windows, WindowWidth, MaxReadLength are mandatory.
  if (c.n_windows < 1 || c.n_windows > MSC_MAX_WINDOWS) return fail("Windows: need 1..32 window offsets");
  if (c.window_width < 1 || c.window_width > MSC_MAX_WINDOW_WIDTH)
    return fail("WindowWidth must be in 1..50 (beyond that the reference itself is undefined: 100 - q2 < WindowWidth at "
                "target position 0, cmd/muscato_screen/main.go:305-313)");
  if (c.max_read_length < 1 || c.max_read_length > MSC_MAX_READ_LENGTH)
    return fail("MaxReadLength must be in 1..1024");
  for (int k = 0; k < c.n_windows; k++)
    if (c.windows[k] < 0) return fail("Windows: negative offset");
  if (c.match_mode != MSC_MATCH_FIRST && c.match_mode != MSC_MATCH_BEST)
    return fail("MatchMode must be 'first' or 'best'");
  // nmiss = int((1 - PMatch) * L) is packed into 11 bits next to the read length: PMatch outside [0, 1]
  // (negative budgets reject everything in the reference, budgets beyond L are meaningless) is refused
  if (!(c.pmatch >= 0.0 && c.pmatch <= 1.0)) return fail("PMatch must be in [0, 1]");
  if (c.max_matches < 1) return fail("MaxMatches must be >= 1");
  if (c.mmtol < 0) return fail("MMTol must be >= 0");

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    (void)cudaGetLastError();
    return fail("no CUDA device available (this library has no CPU fallback)");
  }
  if (c.device < 0 || c.device >= ndev) return fail("device ordinal out of range");
  if (cudaSetDevice(c.device) != cudaSuccess) return fail("cudaSetDevice failed");
  cudaDeviceProp prop{};
  if (cudaGetDeviceProperties(&prop, c.device) != cudaSuccess) return fail("cudaGetDeviceProperties failed");
  if (prop.major != 10) return fail("device is not sm_100 (Blackwell B200); kernels are built for sm_100a only");

  // Random 32-byte sector reads (Bloom words, table buckets, read rows) are the bulk of the traffic
  // at scale; the L2's default fetch granularity pulls whole 128-byte lines from HBM for them.
  // MSC_L2_FETCH = 32 / 64 / 128 sets cudaLimitMaxL2FetchGranularity (a hint; device-wide).
  if (const char* e = getenv("MSC_L2_FETCH")) {
    if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e)) != cudaSuccess) (void)cudaGetLastError();
  }
  msc_ctx* ctx = new msc_ctx();
  ctx->cfg = c;
  ctx->device = c.device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->trace = getenv("MSC_TRACE") && atoi(getenv("MSC_TRACE")) > 0;
  if (const char* e = getenv("MSC_STAGE_EVENTS")) ctx->stage_events = atoi(e) != 0;
  if (const char* e = getenv("MSC_PDL")) {
    ctx->pdl = atoi(e) != 0;
    ctx->pdl_always = atoi(e) == 2;
  }
  ctx->win.nwin = c.n_windows;
  ctx->win.W = c.window_width;
  ctx->win.MRL = c.max_read_length;
  ctx->win.S = (c.max_read_length + 31) / 32;
  ctx->win.min_dinuc = c.min_dinuc;
  for (int k = 0; k < c.n_windows; k++) ctx->win.windows[k] = c.windows[k];
  bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&ctx->ev_rd_free, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&ctx->ev_tg_free, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; ok && i < EV_COUNT; i++) ok = cudaEventCreate(&ctx->ev[i]) == cudaSuccess;
  ok = ok && ctx->counters.reserve(C_COUNT * sizeof(unsigned long long)) == cudaSuccess;
  ok = ok && cudaMallocHost(&ctx->h_counters, C_COUNT * sizeof(unsigned long long)) == cudaSuccess;
  // nmiss table in IEEE double exactly as cmd/muscato_confirm/main.go:198 writes it.
  std::vector<int32_t> nm(c.max_read_length + 1);
  for (int L = 0; L <= c.max_read_length; L++) {
    volatile double one_minus = 1 - c.pmatch;
    volatile double prod = one_minus * (double)L;
    nm[L] = (int32_t)prod;
  }
  {
    // read record word layout (build.cuh, WinCfg): length | budget | sketch | has-X
    int max_nm = 0;
    for (int32_t v : nm) max_nm = std::max(max_nm, (int)v);
    ctx->win.lbits = std::max(1, ceil_log2((uint64_t)c.max_read_length + 1));
    ctx->win.nbits = std::max(1, ceil_log2((uint64_t)max_nm + 1));
    int sk = std::min(12, (31 - ctx->win.lbits - ctx->win.nbits) / 2);
    if (const char* e = getenv("MSC_SKETCH")) sk = std::min(sk, atoi(e));
    ctx->win.sk = sk >= 4 ? sk : 0;
    ctx->win.vm_in_row = 2 * (32 * ctx->win.S - ctx->win.MRL) >= ctx->win.nwin ? 1 : 0;
    if (const char* e = getenv("MSC_VM_IN_ROW")) ctx->win.vm_in_row = ctx->win.vm_in_row && atoi(e) != 0;
  }
  ok = ok && ctx->nmiss.reserve(nm.size() * sizeof(int32_t)) == cudaSuccess;
  ok = ok && cudaMemcpy(ctx->nmiss.p, nm.data(), nm.size() * sizeof(int32_t), cudaMemcpyHostToDevice) == cudaSuccess;
  ok = ok && cudaMemset(ctx->counters.p, 0, C_COUNT * sizeof(unsigned long long)) == cudaSuccess;
  ok = ok && ctx->scan_state.reserve(2 * sizeof(unsigned long long)) == cudaSuccess;
  ok = ok && cudaMemset(ctx->scan_state.p, 0, 2 * sizeof(unsigned long long)) == cudaSuccess;
  // every event is recorded once so that cudaEventElapsedTime never sees a virgin event
  for (int i = 0; ok && i < EV_COUNT; i++) ok = cudaEventRecord(ctx->ev[i], ctx->stream) == cudaSuccess;
  ok = ok && cudaEventRecord(ctx->ev_rd_free, ctx->stream) == cudaSuccess;
  ok = ok && cudaEventRecord(ctx->ev_tg_free, ctx->stream) == cudaSuccess;
  ok = ok && cudaStreamSynchronize(ctx->stream) == cudaSuccess;
  if (!ok) {
    std::string m = std::string("CUDA initialisation failed: ") + cudaGetErrorString(cudaGetLastError());
    msc_destroy(ctx);
    return fail(m.c_str());
  }
  return ctx;
}

void msc_destroy(msc_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  DevBuf* bufs[] = {&ctx->rd_ascii,    &ctx->rd_offs,   &ctx->rd_words, &ctx->rd_x,      &ctx->len_flags, &ctx->validmask, &ctx->rmeta,
                    &ctx->tab,         &ctx->recs,      &ctx->dups,     &ctx->part_count, &ctx->pass_small, &ctx->pass_cnt, &ctx->bloom,
                    &ctx->items,       &ctx->tg_ascii,  &ctx->tg_off,    &ctx->tg_words,
                    &ctx->tg_x,        &ctx->xsum,      &ctx->blk2gene,  &ctx->prep_perm, &ctx->prep_gstart, &ctx->nm_flag, &ctx->nm_pos, &ctx->nm_list, &ctx->cinfo,     &ctx->sizes,     &ctx->pstart,
                    &ctx->block_first, &ctx->match_pre, &ctx->best,     &ctx->rcount,    &ctx->rstart,    &ctx->rfill,
                    &ctx->match_out,   &ctx->long_list, &ctx->mid_list, &ctx->counters,  &ctx->tile_sums, &ctx->scan_state, &ctx->nmiss};
  for (DevBuf* b : bufs) b->release();
  ctx->prep.release_all();
  if (ctx->ev_prep0) cudaEventDestroy(ctx->ev_prep0);
  if (ctx->ev_prep1) cudaEventDestroy(ctx->ev_prep1);
  if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
  for (auto& e : ctx->ev)
    if (e) cudaEventDestroy(e);
  for (auto& te : ctx->trace_ev) cudaEventDestroy(te.second);
  if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
  if (ctx->ev_rd_free) cudaEventDestroy(ctx->ev_rd_free);
  if (ctx->ev_tg_free) cudaEventDestroy(ctx->ev_tg_free);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* msc_last_error(const msc_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

// Entry points that replace inputs or reuse the pair scratch must not run while a deferred
// (MSC_STAGE_DEFER) screen + confirm is still in flight on the stream: the pending run is completed
// and dropped first (its kernels may still be reading the buffers about to be overwritten).
static int drop_deferred(msc_ctx* ctx) {
  if (!ctx->deferred) return MSC_OK;
  ctx->deferred = false;
  RC(sync_counters(ctx));
  ctx->have_cand = ctx->have_confirm = ctx->have_combine = false;
  return MSC_OK;
}

// Size every read-side buffer and the key table for n_reads reads / total ASCII bytes.
static int reads_reserve(msc_ctx* ctx, uint64_t n_reads, uint64_t total) {
  const uint64_t nwin = (uint64_t)ctx->win.nwin;
  // item ids, table slots (+1 in dup_slot) and CSR offsets are 32-bit: at most 2^30 (read, window) items,
  // i.e. a table of at most 2^31 slots, per read set -- larger sets are fed in batches (reads are independent)
  if (n_reads * nwin > (1ull << 30))
    return ctx->fail(MSC_ERR_INPUT, "reads: n_reads * n_windows must be <= 2^30 per read set (feed larger sets in batches)");
  ctx->n_reads = n_reads;
  const int S = ctx->win.S;
  CK(ctx->rd_ascii.reserve(total + 64));
  CK(ctx->rd_offs.reserve((n_reads + 1) * sizeof(uint64_t)));
  CK(ctx->rd_words.reserve((n_reads * S + 2) * sizeof(uint64_t)));
  CK(ctx->rd_x.reserve((n_reads * S + 2) * sizeof(uint64_t)));
  CK(ctx->len_flags.reserve((n_reads + 1) * sizeof(uint32_t)));
  CK(ctx->validmask.reserve((n_reads + 1) * sizeof(uint32_t)));
  CK(ctx->rmeta.reserve((n_reads + 1) * sizeof(uint2)));
  const uint64_t kmax = std::max<uint64_t>(n_reads * nwin, 512);
  // Key table geometry: five-slot buckets at a load factor <= 0.45 of the possible keys, cut into
  // partitions of 2^lg_bpp buckets (16 MB of table each, so that the partition-ordered insert works
  // inside the L2); at most kMaxParts partitions.
  {
    const uint64_t need = (kmax * 4 + 8) / 9;  // buckets: kmax / (5 * 0.45)
    int lg_bpp = 17;
    if (const char* e = getenv("MSC_TABLE_LG_BPP")) lg_bpp = std::min(22, std::max(10, atoi(e)));  // tuning: partition size
    if (need <= (1ull << lg_bpp)) lg_bpp = std::max(3, ceil_log2(need));
    uint64_t parts = (need + (1ull << lg_bpp) - 1) >> lg_bpp;
    while (parts > (uint64_t)kMaxParts) {
      lg_bpp++;
      parts = (need + (1ull << lg_bpp) - 1) >> lg_bpp;
    }
    ctx->tgeo.lg_bpp = lg_bpp;
    ctx->tgeo.n_parts = (uint32_t)parts;
    ctx->tgeo.n_buckets = parts << lg_bpp;
    ctx->n_slots = ctx->tgeo.n_buckets * (uint64_t)kBucketSlots;
    ctx->exact_counts = false;
  }
  // Bloom front sizing (tuning only, never changes results): aim at 32-64 bits per key, but keep
  // the filter L2-resident (<= 64 MB of the 126 MB L2) as long as that still leaves >= 12 bits per
  // key -- a probe served from L2 is ~5x cheaper than one served from an HBM sector, and a 1 %
  // false-positive rate only adds table look-ups.
  if (ctx->cfg.bloom_bits_per_key > 0) {
    ctx->lg_bloom = std::max(10, ceil_log2((kmax * (uint64_t)ctx->cfg.bloom_bits_per_key + 63) / 64));
  } else {
    int lg = std::max(10, ceil_log2((kmax * 32ull + 63) / 64));
    const int lg_l2 = 23;  // 2^23 words = 64 MB
    if (lg > lg_l2) lg = std::max(lg_l2, ceil_log2((kmax * 12ull + 63) / 64));
    ctx->lg_bloom = lg;
  }
  // Exact front (common.cuh, BloomGeom::direct): when one bit per POSSIBLE key (4^W bits) is no larger than the Bloom
  // words would be, the front is that bitmap -- no hash, no false positives -- and the scan tests it slice by slice
  // (<= 32 MB per pass, L2-resident).  W <= 15: 128 MB at most.  MSC_FRONT_DIRECT=0/1 forces the choice,
  // MSC_FRONT_PASS_MB the slice size (tuning only: results do not depend on the front).
  ctx->geom.direct = 0;
  ctx->geom.lg_pass = 0;
  if (ctx->win.W <= 15) {
    const int lg_direct = std::max(10, 2 * ctx->win.W - 6);
    // (measured, profiles/r02/call23_summary.txt: 3.7e7 keys: Bloom 64 MB 3.54 ms vs exact 1.99 ms per 2.5e8 bases;
    // 1.5e7 keys: 1.23 vs 0.79; 6e6 keys, Bloom 32 MB: 0.31 vs 0.36 -- the map pays from a 64 MB Bloom front on)
    bool direct = lg_direct <= ctx->lg_bloom + 1;
    if (const char* e = getenv("MSC_FRONT_DIRECT")) direct = atoi(e) != 0;
    if (direct) {
      int pass_mb = 32;
      if (const char* e = getenv("MSC_FRONT_PASS_MB")) pass_mb = std::max(1, atoi(e));
      const int lg_pass_words = ceil_log2(((uint64_t)pass_mb << 20) / 8);
      ctx->geom.direct = 1;
      ctx->geom.lg_pass = std::max(0, lg_direct - lg_pass_words);
      ctx->lg_bloom = lg_direct;
    }
  }
  // Minimiser geometry of the Bloom addressing (common.cuh): m-mers of the key's first P bases.
  // 4^m is kept >= 4x the number of sectors so that the minimisers spread over all of them, and
  // at most 8 m-mers compete (ALU cost per probed position).  MSC_MINIMIZER_M overrides m.
  {
    const int P = std::min(ctx->win.W, 16);
    // a filter beyond the L2 is addressed in 128-byte lines (what one HBM access delivers), an L2-resident one in
    // 32-byte sectors (what one L1/L2 request moves); MSC_BLOOM_LG_BLK overrides (2 or 4)
    int lg_blk = ctx->lg_bloom > 23 ? 4 : 2;
    if (const char* e = getenv("MSC_BLOOM_LG_BLK")) lg_blk = atoi(e) >= 4 ? 4 : 2;
    lg_blk = std::min(lg_blk, ctx->lg_bloom - 4);
    ctx->geom.lg_blk = lg_blk;
    int m = (ctx->lg_bloom - lg_blk + 2 + 1) / 2;
    if (const char* e = getenv("MSC_MINIMIZER_M")) m = atoi(e);
    m = std::max(m, P - 7);
    m = std::min(std::max(m, 1), P);
    ctx->geom.lg_words = ctx->lg_bloom;
    ctx->geom.m = m;
    ctx->geom.wn = P - m + 1;
    ctx->geom.xr = 0x55555555u & (uint32_t)low_bases_mask(ctx->win.W);
  }
  CK(ctx->tab.reserve(ctx->tgeo.n_buckets * (uint64_t)kBucketBytes));
  CK(ctx->part_count.reserve((size_t)kMaxParts * kPartStride * sizeof(unsigned int)));
  CK(ctx->pass_small.reserve((1ull << ctx->lg_small) * sizeof(uint32_t)));
  CK(ctx->bloom.reserve((1ull << ctx->lg_bloom) * sizeof(uint64_t)));
  CK(ctx->recs.reserve((n_reads * nwin + 1) * sizeof(uint4)));
  CK(ctx->dups.reserve((n_reads * nwin + 1) * sizeof(uint4)));
  CK(ctx->items.reserve((n_reads * nwin + 1) * sizeof(uint2)));
  CK(ctx->best.reserve((n_reads + 1) * sizeof(uint32_t)));
  // rd_words / rd_x rows are read one word past their end by extract32: keep the pad defined.
  CK(cudaMemsetAsync(ctx->rd_words.as<uint64_t>() + n_reads * S, 0, 2 * sizeof(uint64_t), ctx->stream));
  CK(cudaMemsetAsync(ctx->rd_x.as<uint64_t>() + n_reads * S, 0, 2 * sizeof(uint64_t), ctx->stream));
  return MSC_OK;
}

int msc_set_reads(msc_ctx* ctx, const uint8_t* ascii, const uint64_t* offs, uint64_t n_reads) {
  if (!ctx) return MSC_ERR_STATE;
  if (n_reads && (!offs || (!ascii && offs[n_reads] != offs[0]))) return ctx->fail(MSC_ERR_INPUT, "reads: NULL buffer");
  CK(cudaSetDevice(ctx->device));
  RC(drop_deferred(ctx));
  uint64_t total = 0;
  if (n_reads) {
    if (offs[0] != 0) return ctx->fail(MSC_ERR_INPUT, "reads: offs[0] must be 0");
    total = offs[n_reads];
    if (total > n_reads * (uint64_t)ctx->win.MRL)
      return ctx->fail(MSC_ERR_INPUT, "reads: more bytes than n_reads * MaxReadLength (prep_reads truncates, "
                       "cmd/muscato_prep_reads/main.go:67-69)");
  }
  RC(reads_reserve(ctx, n_reads, total));
  ctx->have_reads = false;
  ctx->have_prep = false;
  // Chunked upload: the reads go over in chunks of kUploadReads; the pack and the window pass of a chunk run while the
  // next chunk is on the PCIe link, so that only the partitioned insert is left when the last byte has arrived (the
  // 44 ms pack + build of configs[2] used to start after the whole 10.8 GB copy).  The offsets of a chunk are
  // validated on the host before anything that indexes the ASCII buffer through them is enqueued.
  RC(begin_upload(ctx, ctx->ev_rd_free));
  ctx->st.h2d_bytes += total + (n_reads + 1) * sizeof(uint64_t);
  RC(enqueue_build_begin(ctx, true));
  if (!n_reads) CK(cudaMemsetAsync(ctx->rd_offs.p, 0, sizeof(uint64_t), ctx->copy_stream));
  constexpr uint64_t kUploadReads = 1ull << 22;
  const uint64_t mrl = (uint64_t)ctx->win.MRL;
  for (uint64_t r0 = 0; r0 < n_reads; r0 += kUploadReads) {
    const uint64_t r1 = std::min(n_reads, r0 + kUploadReads);
    uint64_t bad = 0;
    for (uint64_t i = r0; i < r1; i++) bad |= (uint64_t)(offs[i + 1] < offs[i]) | (uint64_t)(offs[i + 1] - offs[i] > mrl);
    bad |= (uint64_t)(offs[r1] > total);
    if (bad) {
      RC(wait_upload_all(ctx));
      for (uint64_t i = r0; i < r1; i++) {
        if (offs[i + 1] < offs[i]) return ctx->fail(MSC_ERR_INPUT, "reads: offsets not monotone at %llu", (unsigned long long)i);
        if (offs[i + 1] - offs[i] > mrl)
          return ctx->fail(MSC_ERR_INPUT, "reads: read %llu is longer than MaxReadLength (prep_reads truncates, "
                           "cmd/muscato_prep_reads/main.go:67-69)", (unsigned long long)i);
      }
      return ctx->fail(MSC_ERR_INPUT, "reads: offsets run past offs[n_reads]");
    }
    // (entry r0 of a later chunk went over as the last entry of the chunk before: the pack kernels may be reading it)
    const uint64_t o0 = r0 ? r0 + 1 : 0;
    CK(cudaMemcpyAsync(ctx->rd_offs.as<uint64_t>() + o0, offs + o0, (r1 - o0 + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->copy_stream));
    if (offs[r1] > offs[r0])
      CK(cudaMemcpyAsync(ctx->rd_ascii.as<uint8_t>() + offs[r0], ascii + offs[r0], offs[r1] - offs[r0], cudaMemcpyHostToDevice, ctx->copy_stream));
    RC(end_upload(ctx));  // the compute stream waits for this chunk only
    RC(enqueue_pack_reads_range(ctx, r0, r1));
    if (r1 == n_reads) CK(cudaEventRecord(ctx->ev_rd_free, ctx->stream));
    RC(enqueue_windows_range(ctx, r0, r1));
  }
  if (!n_reads) {
    RC(end_upload(ctx));
    CK(cudaEventRecord(ctx->ev_rd_free, ctx->stream));
  }
  if (ctx->stage_events) CK(cudaEventRecord(ctx->ev[EV_PACKR1], ctx->stream));
  ctx->have_reads = true;
  RC(enqueue_build_end(ctx, false));  // runs behind the copy; its counters are booked at the next sync
  RC(wait_upload(ctx));          // the caller's buffers are only borrowed for the call
  if (!ctx->cfg.keep_ascii) {
    RC(sync_counters(ctx));
    ctx->rd_ascii.release();
  }
  return MSC_OK;
}

int msc_set_reads_device(msc_ctx* ctx, const uint8_t* d_ascii, const uint64_t* d_offs, uint64_t n_reads,
                         uint64_t total_bytes) {
  if (!ctx) return MSC_ERR_STATE;
  if (n_reads && (!d_offs || (!d_ascii && total_bytes))) return ctx->fail(MSC_ERR_INPUT, "reads: NULL device buffer");
  CK(cudaSetDevice(ctx->device));
  RC(drop_deferred(ctx));
  RC(reads_reserve(ctx, n_reads, total_bytes));
  if (total_bytes) CK(cudaMemcpyAsync(ctx->rd_ascii.p, d_ascii, total_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  if (n_reads) CK(cudaMemcpyAsync(ctx->rd_offs.p, d_offs, (n_reads + 1) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
  else CK(cudaMemsetAsync(ctx->rd_offs.p, 0, sizeof(uint64_t), ctx->stream));
  // the offsets were never seen by the host: validate them on the device
  {
    Filler f;
    f.add(ctx->ctr(C_NLONG), 2 * sizeof(unsigned long long));  // C_PAD1 doubles as the "bad offsets" flag
    RC(enqueue_fill(ctx, f));
  }
  if (n_reads) {
    launch_k(ctx->pdl_on(), validate_read_offsets_kernel, grid_for(n_reads, 256), 256, 0, ctx->stream, 
        ctx->rd_offs.as<uint64_t>(), n_reads, total_bytes, (uint64_t)ctx->win.MRL, ctx->ctr(C_PAD1));
    LAUNCH_CHECK();
  }
  RC(sync_counters(ctx));  // nothing may index the ASCII buffer through unchecked offsets
  if (ctx->h_counters[C_PAD1]) {
    ctx->have_reads = false;
    return ctx->fail(MSC_ERR_INPUT, "reads: device offsets are not monotone, do not end at total_bytes, or a read is longer "
                     "than MaxReadLength");
  }
  ctx->have_reads = true;
  RC(enqueue_build_reads(ctx));
  RC(sync_counters(ctx));  // d_ascii / d_offs are only borrowed for the call
  if (!ctx->cfg.keep_ascii) ctx->rd_ascii.release();
  return MSC_OK;
}

int msc_prep_reads(msc_ctx* ctx, const uint8_t* raw_ascii, const uint64_t* raw_offs, uint64_t n_raw,
                   int32_t min_read_length, uint64_t* n_kept_out, uint64_t* n_unique_out) {
  if (!ctx) return MSC_ERR_STATE;
  if (n_raw && (!raw_offs || (!raw_ascii && raw_offs[n_raw] != raw_offs[0]))) return ctx->fail(MSC_ERR_INPUT, "raw reads: NULL buffer");
  if (n_raw >= 0xffffffffull) return ctx->fail(MSC_ERR_INPUT, "raw reads: more than 2^32-1 reads in one call");
  CK(cudaSetDevice(ctx->device));
  RC(drop_deferred(ctx));
  uint64_t total = 0;
  if (n_raw) {
    if (raw_offs[0] != 0) return ctx->fail(MSC_ERR_INPUT, "raw reads: offs[0] must be 0");
    uint64_t bad = 0;
    for (uint64_t i = 0; i < n_raw; i++) bad |= (uint64_t)(raw_offs[i + 1] < raw_offs[i]);
    if (bad) return ctx->fail(MSC_ERR_INPUT, "raw reads: offsets not monotone");
    total = raw_offs[n_raw];
  }
  ctx->have_prep = false;
  ctx->have_reads = false;
  const int MRL = ctx->win.MRL;
  const int n_planes = (MRL + 1) / 2;
  const uint64_t n = n_raw;
  const uint32_t n_chunks = (uint32_t)std::max<uint64_t>(1, (n + kRadixChunk - 1) / kRadixChunk);
  DevBuf &d_raw = ctx->prep.d_raw, &d_offs = ctx->prep.d_offs, &planes = ctx->prep.planes, &idx_a = ctx->prep.idx_a,
         &idx_b = ctx->prep.idx_b, &keep = ctx->prep.keep, &hist = ctx->prep.hist, &hoff = ctx->prep.hoff,
         &head = ctx->prep.head, &head_scan = ctx->prep.head_scan, &ulen = ctx->prep.ulen, &uoffs = ctx->prep.uoffs;
  auto release_all = [&]() {};  // the scratch stays with the context (msc_destroy releases it)
  if (!ctx->ev_prep0) {
    CK(cudaEventCreate(&ctx->ev_prep0));
    CK(cudaEventCreate(&ctx->ev_prep1));
  }
#define PCK(call)                                                                                          \
  do {                                                                                                     \
    cudaError_t e__ = (call);                                                                              \
    if (e__ != cudaSuccess) {                                                                              \
      release_all();                                                                                       \
      return ctx->fail(MSC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    }                                                                                                      \
  } while (0)
#define PRC(call)           \
  do {                      \
    int rc__ = (call);      \
    if (rc__) {             \
      release_all();        \
      return rc__;          \
    }                       \
  } while (0)
  PCK(d_raw.reserve(total + 64));
  PCK(d_offs.reserve((n + 1) * sizeof(uint64_t)));
  PCK(planes.reserve((size_t)n_planes * std::max<uint64_t>(n, 1)));
  PCK(ctx->prep.key64.reserve((n + 1) * sizeof(uint64_t)));
  PCK(idx_a.reserve((n + 1) * sizeof(uint32_t)));
  PCK(idx_b.reserve((n + 1) * sizeof(uint32_t)));
  PCK(keep.reserve((n + 1) * sizeof(uint32_t)));
  PCK(hist.reserve((size_t)256 * n_chunks * sizeof(uint32_t)));
  PCK(hoff.reserve(((size_t)256 * n_chunks + 1) * sizeof(uint32_t)));
  PCK(head.reserve((n + 1) * sizeof(uint32_t)));
  PCK(head_scan.reserve((n + 2) * sizeof(uint32_t)));
  PCK(ulen.reserve((n + 1) * sizeof(uint32_t)));
  PCK(uoffs.reserve((n + 2) * sizeof(uint64_t)));
  PCK(ctx->prep_perm.reserve((n + 1) * sizeof(uint32_t)));
  PCK(ctx->prep_gstart.reserve((n + 2) * sizeof(uint32_t)));
  if (total) PCK(cudaMemcpyAsync(d_raw.p, raw_ascii, total, cudaMemcpyHostToDevice, ctx->stream));
  if (n) PCK(cudaMemcpyAsync(d_offs.p, raw_offs, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
  ctx->st.h2d_bytes += total + (n + 1) * sizeof(uint64_t);
  PCK(cudaEventRecord(ctx->ev_prep0, ctx->stream));
  {
    Filler f;
    f.add(ctx->ctr(C_PREP_KEPT), 4 * sizeof(unsigned long long));  // C_PREP_KEPT, C_PREP_UNIQ, C_PREP_BYTES, C_PAD3
    f.add(ctx->ctr(C_PREP_TIE), 2 * sizeof(unsigned long long));   // C_PREP_TIE, C_PAD4
    PRC(enqueue_fill(ctx, f));
  }
  uint32_t* cur = idx_a.as<uint32_t>();
  uint32_t* nxt = idx_b.as<uint32_t>();
  if (n) {
    launch_k(ctx->pdl_on(), prep_encode_kernel, grid_for(n, 256), 256, 0, ctx->stream, d_raw.as<uint8_t>(), d_offs.as<uint64_t>(), n, MRL,
                                                                  (int)min_read_length, n_planes, planes.as<uint8_t>(),
                                                                  ctx->prep.key64.as<uint64_t>(), keep.as<uint32_t>(), ctx->ctr(C_PREP_KEPT));
    LAUNCH_CHECK();
    // Stable LSD radix sort of the read indices, one byte plane (two symbols) per pass.  Fast
    // path: sort on the first kPrePlanes planes (16 bases) only and let prep_tiefix_kernel order
    // the short runs of equal prefix by the rest of the key; a run longer than kMaxTieRun (a
    // sequence present in very many copies, or adversarial input) makes the call fall back to
    // passes over all planes.
    constexpr int kPrePlanes = 8;
    auto radix_passes = [&](int first_plane_excl_hi) -> int {  // planes first_plane_excl_hi-1 .. 0
      for (int b = first_plane_excl_hi - 1; b >= 0; b--) {
        const uint8_t* plane = planes.as<uint8_t>() + (size_t)b * n;
        launch_k(ctx->pdl_on(), radix_hist_kernel, n_chunks, kRadixThreads, 0, ctx->stream, cur, plane, n, n_chunks, hist.as<uint32_t>());
        LAUNCH_CHECK();
        RC(enqueue_exclusive_scan<uint32_t>(ctx, hist.as<uint32_t>(), nullptr, (uint64_t)256 * n_chunks, hoff.as<uint32_t>(),
                                            false, ctx->ctr(C_PAD3)));
        launch_k(ctx->pdl_on(), radix_scatter_kernel, n_chunks, kRadixThreads, 0, ctx->stream, cur, plane, n, n_chunks, hoff.as<uint32_t>(), nxt);
        LAUNCH_CHECK();
        std::swap(cur, nxt);
      }
      return MSC_OK;
    };
    bool full_sort = n_planes <= kPrePlanes || (getenv("MSC_PREP_FULL_SORT") && atoi(getenv("MSC_PREP_FULL_SORT")) > 0);
    for (int attempt = 0; attempt < 2; attempt++) {
      launch_k(ctx->pdl_on(), iota_kernel, grid_for(n, 256), 256, 0, ctx->stream, cur, n);
      LAUNCH_CHECK();
      PRC(radix_passes(full_sort ? n_planes : kPrePlanes));
      if (full_sort) break;
      launch_k(ctx->pdl_on(), prep_tiefix_kernel, grid_for(n, 256), 256, 0, ctx->stream, cur, planes.as<uint8_t>(),
               ctx->prep.key64.as<uint64_t>(), n,
               ctx->ctr(C_PREP_KEPT), kPrePlanes, n_planes, ctx->ctr(C_PREP_TIE));
      LAUNCH_CHECK();
      PRC(sync_counters(ctx));  // C_PREP_TIE = longest run that was left unsorted (0: none)
      if (ctx->h_counters[C_PREP_TIE] == 0) break;
      full_sort = true;
    }
    launch_k(ctx->pdl_on(), prep_heads_kernel, grid_for(n, 256), 256, 0, ctx->stream, cur, planes.as<uint8_t>(),
             ctx->prep.key64.as<uint64_t>(), n, ctx->ctr(C_PREP_KEPT), n_planes,
                                                                 head.as<uint32_t>());
    LAUNCH_CHECK();
    PRC(enqueue_exclusive_scan<uint32_t>(ctx, head.as<uint32_t>(), ctx->ctr(C_PREP_KEPT), n, head_scan.as<uint32_t>(), true,
                                         ctx->ctr(C_PREP_UNIQ)));
    launch_k(ctx->pdl_on(), prep_groups_kernel, grid_for(n, 256), 256, 0, ctx->stream, cur, head.as<uint32_t>(), head_scan.as<uint32_t>(),
                                                                  ctx->ctr(C_PREP_KEPT), d_offs.as<uint64_t>(), MRL,
                                                                  ctx->ctr(C_PREP_UNIQ), ctx->prep_gstart.as<uint32_t>(),
                                                                  ulen.as<uint32_t>());
    LAUNCH_CHECK();
    PRC(enqueue_exclusive_scan<uint64_t>(ctx, ulen.as<uint32_t>(), ctx->ctr(C_PREP_UNIQ), n, uoffs.as<uint64_t>(), true,
                                         ctx->ctr(C_PREP_BYTES)));
    PCK(cudaMemcpyAsync(ctx->prep_perm.p, cur, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  PRC(sync_counters(ctx));
  ctx->prep_kept = n ? ctx->h_counters[C_PREP_KEPT] : 0;
  ctx->prep_unique = n ? ctx->h_counters[C_PREP_UNIQ] : 0;
  ctx->prep_bytes = n ? ctx->h_counters[C_PREP_BYTES] : 0;
  const uint64_t U = ctx->prep_unique;
  if (!n || !ctx->prep_kept) PCK(cudaMemsetAsync(ctx->prep_gstart.p, 0, 2 * sizeof(uint32_t), ctx->stream));
  PRC(reads_reserve(ctx, U, ctx->prep_bytes));
  if (U) {
    launch_k(ctx->pdl_on(), prep_gather_kernel, grid_for(U * 32, 256), 256, 0, ctx->stream, d_raw.as<uint8_t>(), d_offs.as<uint64_t>(), ctx->prep_perm.as<uint32_t>(),
                                                                     ctx->prep_gstart.as<uint32_t>(), uoffs.as<uint64_t>(),
                                                                     ctx->ctr(C_PREP_UNIQ), ctx->rd_ascii.as<uint8_t>());
    LAUNCH_CHECK();
    PCK(cudaMemcpyAsync(ctx->rd_offs.p, uoffs.p, (U + 1) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
  } else {
    PCK(cudaMemsetAsync(ctx->rd_offs.p, 0, sizeof(uint64_t), ctx->stream));
  }
  PCK(cudaEventRecord(ctx->ev_prep1, ctx->stream));
  ctx->have_reads = true;
  ctx->have_prep = true;
  PRC(enqueue_build_reads(ctx));
  PRC(sync_counters(ctx));  // the caller's buffers are only borrowed for the call
  {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev_prep0, ctx->ev_prep1) == cudaSuccess) ctx->st.reserved_f[1] += ms;  // "ms_prep"
    else (void)cudaGetLastError();
  }
#undef PCK
#undef PRC
  if (n_kept_out) *n_kept_out = ctx->prep_kept;
  if (n_unique_out) *n_unique_out = U;
  return MSC_OK;
}

int msc_fetch_read_groups(msc_ctx* ctx, uint32_t* perm, uint32_t* group_start) {
  if (!ctx) return MSC_ERR_STATE;
  if (!ctx->have_prep) return ctx->fail(MSC_ERR_STATE, "msc_fetch_read_groups: run msc_prep_reads first");
  CK(cudaSetDevice(ctx->device));
  if (ctx->prep_kept && !perm) return ctx->fail(MSC_ERR_INPUT, "msc_fetch_read_groups: NULL perm");
  if (!group_start) return ctx->fail(MSC_ERR_INPUT, "msc_fetch_read_groups: NULL group_start");
  if (ctx->prep_kept) CK(cudaMemcpyAsync(perm, ctx->prep_perm.p, ctx->prep_kept * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(group_start, ctx->prep_gstart.p, (ctx->prep_unique + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->st.d2h_bytes += ctx->prep_kept * sizeof(uint32_t) + (ctx->prep_unique + 1) * sizeof(uint32_t);
  return MSC_OK;
}

uint64_t msc_unique_reads_bytes(msc_ctx* ctx) { return (ctx && ctx->have_prep) ? ctx->prep_bytes : 0; }

int msc_fetch_unique_reads(msc_ctx* ctx, uint8_t* ascii, uint64_t* offs) {
  if (!ctx) return MSC_ERR_STATE;
  if (!ctx->have_prep) return ctx->fail(MSC_ERR_STATE, "msc_fetch_unique_reads: run msc_prep_reads first");
  if (!ctx->cfg.keep_ascii) return ctx->fail(MSC_ERR_STATE, "msc_fetch_unique_reads needs keep_ascii=1");
  if (!offs || (ctx->prep_bytes && !ascii)) return ctx->fail(MSC_ERR_INPUT, "msc_fetch_unique_reads: NULL buffer");
  CK(cudaSetDevice(ctx->device));
  if (ctx->prep_bytes) CK(cudaMemcpyAsync(ascii, ctx->rd_ascii.p, ctx->prep_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(offs, ctx->rd_offs.p, (ctx->prep_unique + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->st.d2h_bytes += ctx->prep_bytes + (ctx->prep_unique + 1) * sizeof(uint64_t);
  return MSC_OK;
}

int msc_set_targets(msc_ctx* ctx, const uint8_t* ascii, const uint64_t* offs, uint64_t n_targets) {
  if (!ctx) return MSC_ERR_STATE;
  if (n_targets && (!offs || (!ascii && offs[n_targets] != offs[0]))) return ctx->fail(MSC_ERR_INPUT, "targets: NULL buffer");
  CK(cudaSetDevice(ctx->device));
  RC(drop_deferred(ctx));
  uint64_t total = 0;
  std::vector<uint32_t> off32(n_targets + 1, 0);
  if (n_targets) {
    if (offs[0] != 0) return ctx->fail(MSC_ERR_INPUT, "targets: offs[0] must be 0");
    total = offs[n_targets];
    if (total >= 0xffffffffull - 4096ull)
      return ctx->fail(MSC_ERR_INPUT, "targets: more than 2^32-4096 bases in one call; shard the database by target range");
    for (uint64_t i = 0; i <= n_targets; i++) {
      if (i && offs[i] < offs[i - 1]) return ctx->fail(MSC_ERR_INPUT, "targets: offsets not monotone at %llu", (unsigned long long)i);
      off32[i] = (uint32_t)offs[i];
    }
  }
  ctx->n_targets = n_targets;
  ctx->n_bases = total;
  const uint64_t words = (total + 31) / 32;
  ctx->n_tiles = (words + kTileWords - 1) / kTileWords;
  ctx->n_words_alloc = ctx->n_tiles * kTileWords + 64;  // halo + read-past-the-end padding for extract32
  CK(ctx->tg_ascii.reserve(total + 64));
  CK(ctx->tg_off.reserve((n_targets + 2) * sizeof(uint32_t)));
  CK(ctx->tg_words.reserve(ctx->n_words_alloc * sizeof(uint64_t)));
  CK(ctx->tg_x.reserve(ctx->n_words_alloc * sizeof(uint64_t)));
  CK(ctx->xsum.reserve((ctx->n_words_alloc / 32 + 4) * sizeof(uint32_t)));
  // position -> target index at 2^kGeneBlockShift-base granularity (scan kernel, flush_stage)
  const uint64_t n_blk = (total >> kGeneBlockShift) + 2;
  std::vector<uint32_t> blk(n_blk, n_targets ? (uint32_t)(n_targets - 1) : 0u);
  {
    uint64_t g = 0;
    for (uint64_t b = 0; b < n_blk && n_targets; b++) {
      const uint64_t pos = b << kGeneBlockShift;
      while (g + 1 < n_targets && off32[g + 1] <= pos) g++;
      blk[b] = (uint32_t)g;
    }
  }
  CK(ctx->blk2gene.reserve(n_blk * sizeof(uint32_t)));
  RC(begin_upload(ctx, ctx->ev_tg_free));
  CK(cudaMemcpyAsync(ctx->blk2gene.p, blk.data(), n_blk * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->copy_stream));
  if (total) CK(cudaMemcpyAsync(ctx->tg_ascii.p, ascii, total, cudaMemcpyHostToDevice, ctx->copy_stream));
  CK(cudaMemcpyAsync(ctx->tg_off.p, off32.data(), (n_targets + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->copy_stream));
  RC(end_upload(ctx));
  ctx->st.h2d_bytes += total + (n_targets + 1) * sizeof(uint32_t) + n_blk * sizeof(uint32_t);
  ctx->have_targets = true;
  ctx->targets_packed = false;
  RC(enqueue_pack_targets(ctx));
  RC(wait_upload(ctx));  // off32 is a local; the caller's buffers are only borrowed
  if (!ctx->cfg.keep_ascii) {
    RC(sync_counters(ctx));
    ctx->tg_ascii.release();
  }
  return MSC_OK;
}

int msc_set_targets_packed(msc_ctx* ctx, const uint64_t* words, const uint64_t* xplane, const uint64_t* offs, uint64_t n_targets) {
  if (!ctx) return MSC_ERR_STATE;
  if (n_targets && (!offs || (!words && offs[n_targets] != offs[0]))) return ctx->fail(MSC_ERR_INPUT, "packed targets: NULL buffer");
  CK(cudaSetDevice(ctx->device));
  RC(drop_deferred(ctx));
  uint64_t total = 0;
  std::vector<uint32_t> off32(n_targets + 1, 0);
  if (n_targets) {
    if (offs[0] != 0) return ctx->fail(MSC_ERR_INPUT, "targets: offs[0] must be 0");
    total = offs[n_targets];
    if (total >= 0xffffffffull - 4096ull)
      return ctx->fail(MSC_ERR_INPUT, "targets: more than 2^32-4096 bases in one call; shard the database by target range");
    for (uint64_t i = 0; i <= n_targets; i++) {
      if (i && offs[i] < offs[i - 1]) return ctx->fail(MSC_ERR_INPUT, "targets: offsets not monotone at %llu", (unsigned long long)i);
      off32[i] = (uint32_t)offs[i];
    }
  }
  ctx->n_targets = n_targets;
  ctx->n_bases = total;
  const uint64_t words_n = (total + 31) / 32;
  ctx->n_tiles = (words_n + kTileWords - 1) / kTileWords;
  ctx->n_words_alloc = ctx->n_tiles * kTileWords + 64;
  CK(ctx->tg_off.reserve((n_targets + 2) * sizeof(uint32_t)));
  CK(ctx->tg_words.reserve(ctx->n_words_alloc * sizeof(uint64_t)));
  CK(ctx->tg_x.reserve(ctx->n_words_alloc * sizeof(uint64_t)));
  CK(ctx->xsum.reserve((ctx->n_words_alloc / 32 + 4) * sizeof(uint32_t)));
  const uint64_t n_blk = (total >> kGeneBlockShift) + 2;
  std::vector<uint32_t> blk(n_blk, n_targets ? (uint32_t)(n_targets - 1) : 0u);
  {
    uint64_t g = 0;
    for (uint64_t b = 0; b < n_blk && n_targets; b++) {
      const uint64_t pos = b << kGeneBlockShift;
      while (g + 1 < n_targets && off32[g + 1] <= pos) g++;
      blk[b] = (uint32_t)g;
    }
  }
  CK(ctx->blk2gene.reserve(n_blk * sizeof(uint32_t)));
  ctx->tg_ascii.release();  // no ASCII copy exists for packed targets (msc_rebuild(2) is a no-op on them)
  RC(begin_upload(ctx, ctx->ev_tg_free));
  CK(cudaMemcpyAsync(ctx->blk2gene.p, blk.data(), n_blk * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->copy_stream));
  if (words_n) CK(cudaMemcpyAsync(ctx->tg_words.p, words, words_n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->copy_stream));
  if (words_n && xplane) CK(cudaMemcpyAsync(ctx->tg_x.p, xplane, words_n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->copy_stream));
  else if (words_n) CK(cudaMemsetAsync(ctx->tg_x.p, 0, words_n * sizeof(uint64_t), ctx->copy_stream));
  CK(cudaMemcpyAsync(ctx->tg_off.p, off32.data(), (n_targets + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->copy_stream));
  RC(end_upload(ctx));
  ctx->st.h2d_bytes += words_n * sizeof(uint64_t) * (xplane ? 2 : 1) + (n_targets + 1) * sizeof(uint32_t) + n_blk * sizeof(uint32_t);
  ctx->have_targets = true;
  ctx->targets_packed = true;
  if (ctx->stage_events) CK(cudaEventRecord(ctx->ev[EV_PACKT0], ctx->stream));
  {
    Filler f;
    f.add(ctx->xsum.p, (ctx->n_words_alloc / 32 + 4) * sizeof(uint32_t));
    f.add(ctx->ctr(C_TGX), 2 * sizeof(unsigned long long));
    RC(enqueue_fill(ctx, f));
  }
  launch_k(ctx->pdl_on(), packed_targets_finish_kernel, grid_for(ctx->n_words_alloc, 256), 256, 0, ctx->stream, ctx->tg_words.as<uint64_t>(),
           ctx->tg_x.as<uint64_t>(), ctx->n_bases, ctx->n_words_alloc, ctx->xsum.as<uint32_t>(), ctx->ctr(C_TGX));
  LAUNCH_CHECK();
  CK(cudaEventRecord(ctx->ev_tg_free, ctx->stream));
  if (ctx->stage_events) CK(cudaEventRecord(ctx->ev[EV_PACKT1], ctx->stream));
  ctx->have_cand = ctx->have_confirm = ctx->have_combine = false;
  ctx->pend_targets = true;
  RC(wait_upload(ctx));  // off32 / blk are locals; the caller's buffers are only borrowed
  return MSC_OK;
}

// The packed targets of another context on the same device, copied device to device (no host traffic): a second
// context that screens ANOTHER part of the reads against the same database -- read parts are independent of each
// other, so two contexts let the upload of one part run under the scan + confirm of the other (bench.py e2e,
// engine.run_read_parts).
int msc_set_targets_from(msc_ctx* ctx, msc_ctx* src) {
  if (!ctx || !src) return MSC_ERR_STATE;
  if (ctx == src) return MSC_OK;
  if (!src->have_targets) return ctx->fail(MSC_ERR_STATE, "msc_set_targets_from: the source context has no targets");
  if (ctx->device != src->device) return ctx->fail(MSC_ERR_INPUT, "msc_set_targets_from: the contexts are on different devices");
  CK(cudaSetDevice(ctx->device));
  RC(drop_deferred(ctx));
  ctx->n_targets = src->n_targets;
  ctx->n_bases = src->n_bases;
  ctx->n_tiles = src->n_tiles;
  ctx->n_words_alloc = src->n_words_alloc;
  const uint64_t n_blk = (ctx->n_bases >> kGeneBlockShift) + 2;
  const size_t xsum_bytes = (ctx->n_words_alloc / 32 + 4) * sizeof(uint32_t);
  CK(ctx->tg_off.reserve((ctx->n_targets + 2) * sizeof(uint32_t)));
  CK(ctx->tg_words.reserve(ctx->n_words_alloc * sizeof(uint64_t)));
  CK(ctx->tg_x.reserve(ctx->n_words_alloc * sizeof(uint64_t)));
  CK(ctx->xsum.reserve(xsum_bytes));
  CK(ctx->blk2gene.reserve(n_blk * sizeof(uint32_t)));
  ctx->tg_ascii.release();  // as for msc_set_targets_packed: msc_rebuild(2) has nothing to redo
  // the source's pack kernel has finished when its "ASCII buffer is free" event has
  CK(cudaStreamWaitEvent(ctx->stream, src->ev_tg_free, 0));
  if (ctx->stage_events) CK(cudaEventRecord(ctx->ev[EV_PACKT0], ctx->stream));
  CK(cudaMemcpyAsync(ctx->tg_off.p, src->tg_off.p, (ctx->n_targets + 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->tg_words.p, src->tg_words.p, ctx->n_words_alloc * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->tg_x.p, src->tg_x.p, ctx->n_words_alloc * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->xsum.p, src->xsum.p, xsum_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->blk2gene.p, src->blk2gene.p, n_blk * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->ctr(C_TGX), src->ctr(C_TGX), 2 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaEventRecord(ctx->ev_tg_free, ctx->stream));
  if (ctx->stage_events) CK(cudaEventRecord(ctx->ev[EV_PACKT1], ctx->stream));
  ctx->have_targets = true;
  ctx->targets_packed = true;
  ctx->have_cand = ctx->have_confirm = ctx->have_combine = false;
  ctx->pend_targets = true;
  CK(cudaStreamSynchronize(ctx->stream));  // the source may replace its targets as soon as this call has returned
  return MSC_OK;
}

uint64_t msc_packed_target_words(const msc_ctx* ctx) { return (ctx && ctx->have_targets) ? (ctx->n_bases + 31) / 32 : 0; }

int msc_fetch_packed_targets(msc_ctx* ctx, uint64_t* words, uint64_t* xplane, int32_t* has_x) {
  if (!ctx) return MSC_ERR_STATE;
  if (!ctx->have_targets) return ctx->fail(MSC_ERR_STATE, "msc_fetch_packed_targets: no targets set");
  CK(cudaSetDevice(ctx->device));
  RC(drop_deferred(ctx));
  RC(sync_counters(ctx));
  const uint64_t n = (ctx->n_bases + 31) / 32;
  if (n && (!words || !xplane)) return ctx->fail(MSC_ERR_INPUT, "msc_fetch_packed_targets: NULL buffer");
  if (n) {
    CK(cudaMemcpyAsync(words, ctx->tg_words.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(xplane, ctx->tg_x.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->st.d2h_bytes += 2 * n * sizeof(uint64_t);
  }
  if (has_x) *has_x = ctx->h_counters[C_TGX] ? 1 : 0;
  return MSC_OK;
}

int msc_rebuild(msc_ctx* ctx, int what) {
  if (!ctx) return MSC_ERR_STATE;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->cfg.keep_ascii) return ctx->fail(MSC_ERR_STATE, "msc_rebuild needs keep_ascii=1");
  if ((what & 1) && !ctx->have_reads) return ctx->fail(MSC_ERR_STATE, "msc_rebuild: no reads set");
  if ((what & 2) && !ctx->have_targets) return ctx->fail(MSC_ERR_STATE, "msc_rebuild: no targets set");
  return run_pipeline(ctx, what & 3, false, false, false);
}

int msc_screen(msc_ctx* ctx) {
  if (!ctx) return MSC_ERR_STATE;
  if (!ctx->have_reads || !ctx->have_targets) return ctx->fail(MSC_ERR_STATE, "msc_screen: set reads and targets first");
  CK(cudaSetDevice(ctx->device));
  return run_pipeline(ctx, 0, true, false, false);
}

int msc_confirm(msc_ctx* ctx) {
  if (!ctx) return MSC_ERR_STATE;
  if (!ctx->have_cand) return ctx->fail(MSC_ERR_STATE, "msc_confirm: run msc_screen first");
  CK(cudaSetDevice(ctx->device));
  return run_pipeline(ctx, 0, false, true, false);
}

void* msc_best_device(msc_ctx* ctx) { return (ctx && (ctx->have_confirm || ctx->deferred)) ? ctx->best.p : nullptr; }

void* msc_stream(msc_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

void* msc_matches_device(msc_ctx* ctx, uint64_t* n) {
  if (!ctx || !ctx->have_combine) return nullptr;
  if (n) *n = ctx->n_match;
  return ctx->match_out.p;
}

int msc_combine(msc_ctx* ctx) {
  if (!ctx) return MSC_ERR_STATE;
  if (!ctx->have_confirm) return ctx->fail(MSC_ERR_STATE, "msc_combine: run msc_confirm first");
  CK(cudaSetDevice(ctx->device));
  if (ctx->n_match_pre >= 0xffffffffull) return ctx->fail(MSC_ERR_NOMEM, "more than 2^32 matches in one batch");
  return run_pipeline(ctx, 0, false, false, true);
}

int msc_run(msc_ctx* ctx) {
  if (!ctx) return MSC_ERR_STATE;
  if (!ctx->have_reads || !ctx->have_targets) return ctx->fail(MSC_ERR_STATE, "msc_run: set reads and targets first");
  CK(cudaSetDevice(ctx->device));
  return run_pipeline(ctx, 0, true, true, true);
}

int msc_run_stages(msc_ctx* ctx, int rebuild_what, int stages) {
  if (!ctx) return MSC_ERR_STATE;
  if (!ctx->have_reads || !ctx->have_targets) return ctx->fail(MSC_ERR_STATE, "msc_run_stages: set reads and targets first");
  if ((rebuild_what & 3) && !ctx->cfg.keep_ascii) return ctx->fail(MSC_ERR_STATE, "rebuild needs keep_ascii=1");
  const bool scan = stages & MSC_STAGE_SCREEN, conf = stages & MSC_STAGE_CONFIRM, comb = stages & MSC_STAGE_COMBINE;
  const bool defer = stages & MSC_STAGE_DEFER;
  if (defer && (!scan || !conf || comb))
    return ctx->fail(MSC_ERR_STATE, "MSC_STAGE_DEFER goes with SCREEN | CONFIRM (combine follows in its own call)");
  if (conf && !scan && (!ctx->have_cand || (rebuild_what & 3)))
    return ctx->fail(MSC_ERR_STATE, "msc_run_stages: confirm needs candidates (run the screen stage)");
  if (comb && !conf && !ctx->deferred && (!ctx->have_confirm || scan || (rebuild_what & 3)))
    return ctx->fail(MSC_ERR_STATE, "msc_run_stages: combine needs confirmed pairs");
  CK(cudaSetDevice(ctx->device));
  return run_pipeline(ctx, rebuild_what & 3, scan, conf, comb, defer);
}

int msc_rebuild_and_run(msc_ctx* ctx, int what) {
  return msc_run_stages(ctx, what, MSC_STAGE_SCREEN | MSC_STAGE_CONFIRM | MSC_STAGE_COMBINE);
}

int msc_set_shards(msc_ctx* ctx, int32_t n_shards) {
  if (!ctx) return MSC_ERR_STATE;
  if (n_shards < 1) return ctx->fail(MSC_ERR_CONFIG, "msc_set_shards: n_shards must be >= 1");
  ctx->n_shards = n_shards;
  ctx->shard_overflow = false;
  return MSC_OK;
}

int msc_shard_overflow(const msc_ctx* ctx) { return ctx && ctx->shard_overflow ? 1 : 0; }

int msc_overflow_keys(msc_ctx* ctx, uint64_t** keys, uint64_t* n) {
  if (!ctx || !keys || !n) return MSC_ERR_STATE;
  if (!ctx->have_confirm || ctx->deferred) return ctx->fail(MSC_ERR_STATE, "msc_overflow_keys: run msc_confirm first");
  CK(cudaSetDevice(ctx->device));
  std::vector<uint64_t> fps;
  RC(shard_overflow_keys(ctx, fps));
  *n = fps.size();
  *keys = (uint64_t*)malloc(std::max<size_t>(fps.size(), 1) * sizeof(uint64_t));
  if (!*keys) return ctx->fail(MSC_ERR_NOMEM, "host allocation failed");
  if (!fps.empty()) memcpy(*keys, fps.data(), fps.size() * sizeof(uint64_t));
  return MSC_OK;
}

uint32_t msc_diverted_record_bytes(const msc_ctx* ctx) { return ctx ? div_rec_bytes(ctx) : 0; }

int msc_divert_groups(msc_ctx* ctx, const uint64_t* keys, uint64_t n_keys, uint32_t gene_base, uint8_t** recs, uint64_t* n_recs) {
  if (!ctx || !recs || !n_recs || (n_keys && !keys)) return MSC_ERR_STATE;
  if (!ctx->have_cand || !ctx->have_confirm || ctx->deferred)
    return ctx->fail(MSC_ERR_STATE, "msc_divert_groups: run msc_screen and msc_confirm first");
  CK(cudaSetDevice(ctx->device));
  std::vector<uint8_t> buf;
  uint64_t nr = 0;
  RC(shard_divert(ctx, keys, n_keys, gene_base, buf, nr));
  *n_recs = nr;
  *recs = (uint8_t*)malloc(std::max<size_t>(buf.size(), 1));
  if (!*recs) return ctx->fail(MSC_ERR_NOMEM, "host allocation failed");
  if (!buf.empty()) memcpy(*recs, buf.data(), buf.size());
  return MSC_OK;
}

int msc_replay_diverted(msc_ctx* ctx, const uint8_t* recs, uint64_t n_recs, msc_match** out, uint64_t* n) {
  if (!ctx || !out || !n || (n_recs && !recs)) return MSC_ERR_STATE;
  if (!ctx->have_reads) return ctx->fail(MSC_ERR_STATE, "msc_replay_diverted: no reads set");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  std::vector<uint4> surv;
  RC(shard_replay(ctx, recs, n_recs, surv));
  *n = surv.size();
  *out = (msc_match*)malloc(std::max<size_t>(surv.size(), 1) * sizeof(msc_match));
  if (!*out) return ctx->fail(MSC_ERR_NOMEM, "host allocation failed");
  if (!surv.empty()) memcpy(*out, surv.data(), surv.size() * sizeof(msc_match));
  return MSC_OK;
}

int msc_fetch_matches_into(msc_ctx* ctx, msc_match* dst, uint64_t capacity, uint64_t* n) {
  if (!ctx || !n) return MSC_ERR_STATE;
  if (!ctx->have_combine) return ctx->fail(MSC_ERR_STATE, "msc_fetch_matches: run msc_combine first");
  CK(cudaSetDevice(ctx->device));
  *n = ctx->n_match;
  if (ctx->n_match > capacity)
    return ctx->fail(MSC_ERR_NOMEM, "msc_fetch_matches_into: %llu matches do not fit %llu slots",
                     (unsigned long long)ctx->n_match, (unsigned long long)capacity);
  if (ctx->n_match) {
    if (!dst) return ctx->fail(MSC_ERR_INPUT, "msc_fetch_matches_into: NULL destination");
    static_assert(sizeof(msc_match) == sizeof(uint4), "msc_match layout");
    CK(cudaMemcpyAsync(dst, ctx->match_out.p, ctx->n_match * sizeof(msc_match), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->st.d2h_bytes += ctx->n_match * sizeof(msc_match);
  }
  return MSC_OK;
}

int msc_fetch_nonmatch(msc_ctx* ctx, uint32_t* ids, uint64_t capacity, uint64_t* n) {
  if (!ctx || !n) return MSC_ERR_STATE;
  if (!ctx->have_confirm) return ctx->fail(MSC_ERR_STATE, "msc_fetch_nonmatch: run msc_confirm first");
  CK(cudaSetDevice(ctx->device));
  const uint64_t U = ctx->n_reads;
  *n = 0;
  if (!U) return MSC_OK;
  CK(ctx->nm_flag.reserve((U + 4) * sizeof(uint32_t)));
  CK(ctx->nm_pos.reserve((U + 4) * sizeof(uint32_t)));
  CK(ctx->nm_list.reserve((U + 4) * sizeof(uint32_t)));
  launch_k(ctx->pdl_on(), nonmatch_flag_kernel, grid_for(U, 256), 256, 0, ctx->stream, ctx->best.as<uint32_t>(), U, MSC_NO_MATCH,
                                                                   ctx->nm_flag.as<uint32_t>());
  LAUNCH_CHECK();
  RC(enqueue_exclusive_scan<uint32_t>(ctx, ctx->nm_flag.as<uint32_t>(), nullptr, U, ctx->nm_pos.as<uint32_t>(), true,
                                      ctx->ctr(C_PAD2)));
  launch_k(ctx->pdl_on(), nonmatch_scatter_kernel, grid_for(U, 256), 256, 0, ctx->stream, ctx->nm_flag.as<uint32_t>(), ctx->nm_pos.as<uint32_t>(), U,
                                                                      ctx->nm_list.as<uint32_t>());
  LAUNCH_CHECK();
  RC(sync_counters(ctx));
  const uint64_t cnt = ctx->h_counters[C_PAD2];
  *n = cnt;
  if (!ids) return MSC_OK;
  if (cnt > capacity)
    return ctx->fail(MSC_ERR_NOMEM, "msc_fetch_nonmatch: %llu reads do not fit %llu slots", (unsigned long long)cnt,
                     (unsigned long long)capacity);
  if (cnt) {
    CK(cudaMemcpyAsync(ids, ctx->nm_list.p, cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->st.d2h_bytes += cnt * sizeof(uint32_t);
  }
  return MSC_OK;
}

int msc_fetch_matches(msc_ctx* ctx, msc_match** out, uint64_t* n) {
  if (!ctx || !out || !n) return MSC_ERR_STATE;
  if (!ctx->have_combine) return ctx->fail(MSC_ERR_STATE, "msc_fetch_matches: run msc_combine first");
  msc_match* h = (msc_match*)malloc(std::max<uint64_t>(1, ctx->n_match) * sizeof(msc_match));
  if (!h) return ctx->fail(MSC_ERR_NOMEM, "host allocation failed");
  const int rc = msc_fetch_matches_into(ctx, h, std::max<uint64_t>(1, ctx->n_match), n);
  if (rc) {
    free(h);
    return rc;
  }
  *out = h;
  return MSC_OK;
}

int msc_set_stage_timing(msc_ctx* ctx, int on) {
  if (!ctx) return MSC_ERR_STATE;
  CK(cudaSetDevice(ctx->device));
  if (ctx->pend_reads || ctx->pend_targets) RC(sync_counters(ctx));  // book builds enqueued under the old setting
  ctx->stage_events = on != 0;
  return MSC_OK;
}

int msc_get_stats(const msc_ctx* cctx, msc_stats* out) {
  if (!cctx || !out) return MSC_ERR_STATE;
  msc_ctx* ctx = const_cast<msc_ctx*>(cctx);
  if (ctx->pend_reads || ctx->pend_targets) {  // a build is still in flight: book it first
    CK(cudaSetDevice(ctx->device));
    RC(sync_counters(ctx));
  }
  *out = ctx->st;
  return MSC_OK;
}

void msc_reset_stats(msc_ctx* ctx) {
  if (!ctx) return;
  msc_stats z{};
  z.n_reads = ctx->st.n_reads;
  z.n_keys = ctx->st.n_keys;
  z.n_key_groups = ctx->st.n_key_groups;
  z.table_slots = ctx->st.table_slots;
  z.bloom_bytes = ctx->st.bloom_bytes;
  z.n_targets = ctx->st.n_targets;
  z.target_bases = ctx->st.target_bases;
  ctx->st = z;
}

void msc_free(void* p) { free(p); }

int msc_dump_keys(msc_ctx* ctx, msc_key_rec** out, uint64_t* n) {
  if (!ctx || !out || !n) return MSC_ERR_STATE;
  if (!ctx->have_reads) return ctx->fail(MSC_ERR_STATE, "msc_dump_keys: no reads set");
  CK(cudaSetDevice(ctx->device));
  RC(sync_counters(ctx));
  // group members: slot-resident first items + the CSR of further members
  std::vector<uint32_t> items;
  {
    const uint64_t nb = ctx->tgeo.n_buckets;
    const uint64_t chunk = 1ull << 18;  // buckets per D2H chunk (32 MB)
    std::vector<uint8_t> buf(chunk * kBucketBytes);
    for (uint64_t b0 = 0; b0 < nb; b0 += chunk) {
      const uint64_t n = std::min(chunk, nb - b0);
      CK(cudaMemcpy(buf.data(), ctx->tab.as<uint8_t>() + b0 * kBucketBytes, n * kBucketBytes, cudaMemcpyDeviceToHost));
      for (uint64_t b = 0; b < n; b++) {
        const uint8_t* bp = buf.data() + b * kBucketBytes;
        for (int sl = 0; sl < kBucketSlots; sl++) {
          uint64_t fp;
          memcpy(&fp, bp + 8 * sl, 8);
          if (!fp) continue;
          uint32_t item0;
          memcpy(&item0, bp + kBucketRecOff + 16 * sl, 4);
          items.push_back(item0);
        }
      }
    }
    std::vector<uint2> dupv(ctx->n_dup);  // CSR entries: (item, read record word)
    if (ctx->n_dup) CK(cudaMemcpy(dupv.data(), ctx->items.p, ctx->n_dup * sizeof(uint2), cudaMemcpyDeviceToHost));
    for (const uint2& e : dupv) items.push_back(e.x);
  }
  if (items.size() != ctx->n_keys)
    return ctx->fail(MSC_ERR_STATE, "key table inconsistent: %llu items for %llu keys", (unsigned long long)items.size(),
                     (unsigned long long)ctx->n_keys);
  msc_key_rec* h = (msc_key_rec*)malloc(std::max<uint64_t>(1, ctx->n_keys) * sizeof(msc_key_rec));
  if (!h) return ctx->fail(MSC_ERR_NOMEM, "host allocation failed");
  const uint32_t nwin = (uint32_t)ctx->win.nwin;
  for (uint64_t i = 0; i < ctx->n_keys; i++) {
    h[i].read_id = items[i] / nwin;
    h[i].window = items[i] % nwin;
  }
  std::sort(h, h + ctx->n_keys, [](const msc_key_rec& a, const msc_key_rec& b) {
    return a.window != b.window ? a.window < b.window : a.read_id < b.read_id;
  });
  *out = h;
  *n = ctx->n_keys;
  return MSC_OK;
}

int msc_dump_candidates(msc_ctx* ctx, msc_cand_rec** out, uint64_t* n) {
  if (!ctx || !out || !n) return MSC_ERR_STATE;
  if (!ctx->have_cand) return ctx->fail(MSC_ERR_STATE, "msc_dump_candidates: run msc_screen first");
  CK(cudaSetDevice(ctx->device));
  RC(drop_deferred(ctx));
  ScopedDevBuf tmp;
  uint64_t cnt = 0;
  int rc = MSC_OK;
  for (int attempt = 0; attempt < 5; attempt++) {
    rc = mark_expand_start(ctx);
    if (rc == MSC_OK) rc = enqueue_pairs(ctx, 1, tmp);
    if (rc == MSC_OK) rc = sync_counters(ctx);
    if (rc == MSC_OK) rc = finish_pairs(ctx, tmp, &cnt);
    if (rc != NEED_RETRY) break;
  }
  ctx->have_confirm = ctx->have_combine = false;  // the pair kernel scratch (best / pass counts) was reused
  if (rc == NEED_RETRY) rc = ctx->fail(MSC_ERR_NOMEM, "candidate dump buffer kept overflowing");
  if (rc) {
    tmp.release();
    return rc;
  }
  msc_cand_rec* h = (msc_cand_rec*)malloc(std::max<uint64_t>(1, cnt) * sizeof(msc_cand_rec));
  if (!h) {
    tmp.release();
    return ctx->fail(MSC_ERR_NOMEM, "host allocation failed");
  }
  static_assert(sizeof(msc_cand_rec) == sizeof(uint4), "msc_cand_rec layout");
  cudaError_t e = cnt ? cudaMemcpy(h, tmp.p, cnt * sizeof(uint4), cudaMemcpyDeviceToHost) : cudaSuccess;
  tmp.release();
  if (e != cudaSuccess) {
    free(h);
    return ctx->fail(MSC_ERR_CUDA, "D2H failed: %s", cudaGetErrorString(e));
  }
  std::sort(h, h + cnt, [](const msc_cand_rec& a, const msc_cand_rec& b) {
    if (a.window != b.window) return a.window < b.window;
    if (a.gene_id != b.gene_id) return a.gene_id < b.gene_id;
    if (a.p != b.p) return a.p < b.p;
    return a.read_id < b.read_id;
  });
  *out = h;
  *n = cnt;
  return MSC_OK;
}

}  // extern "C"
