// api.cu -- context management and the C ABI of include/muscato_b200.h.
// One context = one CUDA device + one stream.  There is deliberately no CPU fallback:
// every entry point that computes anything launches the sm_100a kernels in this directory
// and fails with MSC_ERR_CUDA when that is impossible.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/muscato_b200.h"
#include "build.cuh"
#include "combine.cuh"
#include "common.cuh"
#include "confirm.cuh"
#include "pack.cuh"
#include "prefix.cuh"
#include "scan.cuh"

using namespace msc;

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
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

enum Counter { C_NKEYS = 0, C_NGROUPS, C_NCAND, C_BLOOMPASS, C_NMATCH, C_NPASS, C_SCANTOTAL, C_NOVER, C_COUNT };

}  // namespace

struct msc_ctx {
  msc_config cfg{};
  WinCfg win{};
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  std::string err;
  msc_stats st{};

  // reads
  uint64_t n_reads = 0, rd_ascii_bytes = 0;
  bool have_reads = false;
  DevBuf rd_ascii, rd_offs, rd_words, rd_x, len_flags, validmask;
  // key table
  int lg_slots = 0, lg_bloom = 0;
  uint64_t n_keys = 0, n_groups = 0;
  DevBuf tab_fp, tab_cnt, tab_start, tab_fill, bloom, items;
  // targets
  uint64_t n_targets = 0, n_bases = 0, n_words_alloc = 0, n_tiles = 0;
  bool have_targets = false;
  DevBuf tg_ascii, tg_off, tg_words, tg_x, xsum;
  // candidates / pairs
  uint64_t n_cand = 0, n_pairs = 0;
  bool have_cand = false;
  DevBuf cand, cinfo, sizes, pstart, block_first;
  // matches
  uint64_t n_match_pre = 0, n_match = 0;
  bool have_confirm = false, have_combine = false;
  DevBuf match_pre, best, rcount, rstart, rfill, match_out;
  // misc
  DevBuf counters, tile_sums, nmiss;
  unsigned long long* h_counters = nullptr;  // pinned mirror

  int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
};

#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return ctx->fail(MSC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

#define LAUNCH_CHECK()                 \
  do {                                 \
    ctx->st.kernel_launches++;         \
    CK(cudaGetLastError());            \
  } while (0)

namespace {

inline unsigned grid_for(uint64_t n, int block) { return (unsigned)((n + block - 1) / block); }

int ceil_log2(uint64_t v) {
  int l = 0;
  while ((1ull << l) < v) l++;
  return l;
}

// Exclusive scan of uint32 in[n] -> OutT out[n] (+ out[n] = total when write_end).
// The grand total is also left in counters[C_SCANTOTAL].
template <typename OutT>
int device_exclusive_scan(msc_ctx* ctx, const uint32_t* in, uint64_t n, OutT* out, bool write_end) {
  unsigned long long* total = ctx->counters.as<unsigned long long>() + C_SCANTOTAL;
  if (n == 0) {
    CK(cudaMemsetAsync(total, 0, sizeof(unsigned long long), ctx->stream));
    if (write_end) CK(cudaMemsetAsync(out, 0, sizeof(OutT), ctx->stream));
    return MSC_OK;
  }
  const uint64_t ntiles = (n + kScanTile - 1) / kScanTile;
  CK(ctx->tile_sums.reserve(ntiles * sizeof(uint64_t)));
  scan_tile_sums<<<(unsigned)ntiles, kScanThreads, 0, ctx->stream>>>(in, n, ctx->tile_sums.as<uint64_t>());
  LAUNCH_CHECK();
  scan_tile_offsets<<<1, kScanThreads, 0, ctx->stream>>>(ctx->tile_sums.as<uint64_t>(), ntiles,
                                                         reinterpret_cast<uint64_t*>(total));
  LAUNCH_CHECK();
  scan_apply<OutT><<<(unsigned)ntiles, kScanThreads, 0, ctx->stream>>>(in, n, ctx->tile_sums.as<uint64_t>(), out,
                                                                       write_end ? 1 : 0);
  LAUNCH_CHECK();
  return MSC_OK;
}

int fetch_counters(msc_ctx* ctx) {
  CK(cudaMemcpyAsync(ctx->h_counters, ctx->counters.p, C_COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                     ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->st.d2h_bytes += C_COUNT * sizeof(unsigned long long);
  return MSC_OK;
}

int zero_counter(msc_ctx* ctx, int which) {
  CK(cudaMemsetAsync(ctx->counters.as<unsigned long long>() + which, 0, sizeof(unsigned long long), ctx->stream));
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

// ---- reads: device-side pack + key table build (from the resident ASCII copy) -------------
int build_reads_device(msc_ctx* ctx) {
  const uint64_t U = ctx->n_reads;
  const int S = ctx->win.S;
  const uint64_t nw = U * (uint64_t)S;
  CK(cudaEventRecord(ctx->ev[0], ctx->stream));
  CK(cudaMemsetAsync(ctx->len_flags.p, 0, (U + 1) * sizeof(uint32_t), ctx->stream));
  if (nw) {
    pack_reads_kernel<<<grid_for(nw, 256), 256, 0, ctx->stream>>>(
        ctx->rd_ascii.as<uint8_t>(), ctx->rd_offs.as<uint64_t>(), U, S, ctx->rd_words.as<uint64_t>(),
        ctx->rd_x.as<uint64_t>(), ctx->len_flags.as<uint32_t>());
    LAUNCH_CHECK();
  }
  CK(cudaEventRecord(ctx->ev[1], ctx->stream));

  const uint64_t slots = 1ull << ctx->lg_slots;
  const uint64_t bwords = 1ull << ctx->lg_bloom;
  CK(cudaMemsetAsync(ctx->tab_fp.p, 0, slots * sizeof(uint64_t), ctx->stream));
  CK(cudaMemsetAsync(ctx->tab_cnt.p, 0, slots * sizeof(uint32_t), ctx->stream));
  CK(cudaMemsetAsync(ctx->tab_fill.p, 0, slots * sizeof(uint32_t), ctx->stream));
  CK(cudaMemsetAsync(ctx->bloom.p, 0, bwords * sizeof(uint64_t), ctx->stream));
  CK(cudaMemsetAsync(ctx->validmask.p, 0, (U + 1) * sizeof(uint32_t), ctx->stream));
  if (int rc = zero_counter(ctx, C_NKEYS)) return rc;
  if (int rc = zero_counter(ctx, C_NGROUPS)) return rc;
  if (U) {
    BuildArgs a{};
    a.rd_words = ctx->rd_words.as<uint64_t>();
    a.rd_x = ctx->rd_x.as<uint64_t>();
    a.len_flags = ctx->len_flags.as<uint32_t>();
    a.n_reads = U;
    a.tab_fp = ctx->tab_fp.as<uint64_t>();
    a.tab_cnt = ctx->tab_cnt.as<uint32_t>();
    a.lg_slots = ctx->lg_slots;
    a.bloom = ctx->bloom.as<unsigned long long>();
    a.lg_bloom = ctx->lg_bloom;
    a.validmask = ctx->validmask.as<uint32_t>();
    a.n_keys = ctx->counters.as<unsigned long long>() + C_NKEYS;
    a.n_groups = ctx->counters.as<unsigned long long>() + C_NGROUPS;
    build_insert_kernel<<<grid_for(U, 256), 256, 0, ctx->stream>>>(ctx->win, a);
    LAUNCH_CHECK();
  }
  if (int rc = device_exclusive_scan<uint32_t>(ctx, ctx->tab_cnt.as<uint32_t>(), slots, ctx->tab_start.as<uint32_t>(),
                                               true))
    return rc;
  if (U) {
    build_fill_kernel<<<grid_for(U, 256), 256, 0, ctx->stream>>>(
        ctx->win, ctx->rd_words.as<uint64_t>(), ctx->rd_x.as<uint64_t>(), ctx->len_flags.as<uint32_t>(),
        ctx->validmask.as<uint32_t>(), U, ctx->tab_fp.as<uint64_t>(), ctx->tab_start.as<uint32_t>(),
        ctx->tab_fill.as<uint32_t>(), ctx->lg_slots, ctx->items.as<uint32_t>());
    LAUNCH_CHECK();
  }
  CK(cudaEventRecord(ctx->ev[2], ctx->stream));
  if (int rc = fetch_counters(ctx)) return rc;
  ctx->n_keys = ctx->h_counters[C_NKEYS];
  ctx->n_groups = ctx->h_counters[C_NGROUPS];
  ctx->st.n_reads = U;
  ctx->st.n_keys = ctx->n_keys;
  ctx->st.n_key_groups = ctx->n_groups;
  ctx->st.table_slots = slots;
  ctx->st.bloom_bytes = bwords * sizeof(uint64_t);
  ctx->st.ms_pack_reads += elapsed(ctx, 0, 1);
  ctx->st.ms_build += elapsed(ctx, 1, 2);
  ctx->have_cand = ctx->have_confirm = ctx->have_combine = false;
  return MSC_OK;
}

int pack_targets_device(msc_ctx* ctx) {
  CK(cudaEventRecord(ctx->ev[0], ctx->stream));
  CK(cudaMemsetAsync(ctx->tg_x.p, 0, ctx->n_words_alloc * sizeof(uint64_t), ctx->stream));
  CK(cudaMemsetAsync(ctx->xsum.p, 0, (ctx->n_words_alloc / 32 + 4) * sizeof(uint32_t), ctx->stream));
  pack_targets_kernel<<<grid_for(ctx->n_words_alloc, 256), 256, 0, ctx->stream>>>(
      ctx->tg_ascii.as<uint8_t>(), ctx->n_bases, ctx->tg_words.as<uint64_t>(), ctx->n_words_alloc,
      ctx->tg_x.as<uint64_t>(), ctx->xsum.as<uint32_t>());
  LAUNCH_CHECK();
  CK(cudaEventRecord(ctx->ev[1], ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->st.ms_pack_targets += elapsed(ctx, 0, 1);
  ctx->st.n_targets = ctx->n_targets;
  ctx->st.target_bases = ctx->n_bases;
  ctx->have_cand = ctx->have_confirm = ctx->have_combine = false;
  return MSC_OK;
}

int run_confirm_kernel(msc_ctx* ctx, int mode, DevBuf& outbuf, uint64_t* n_out) {
  // Runs the pair kernel; grows outbuf and re-runs if the output did not fit.
  for (int attempt = 0; attempt < 3; attempt++) {
    const uint64_t cap = outbuf.cap / sizeof(uint4);
    if (int rc = zero_counter(ctx, C_NMATCH)) return rc;
    if (int rc = zero_counter(ctx, C_NPASS)) return rc;
    CK(cudaMemsetAsync(ctx->best.p, 0x7f, (ctx->n_reads + 1) * sizeof(uint32_t), ctx->stream));  // MSC_NO_MATCH
    CK(cudaMemsetAsync(ctx->tab_fill.p, 0, (1ull << ctx->lg_slots) * sizeof(uint32_t), ctx->stream));
    if (ctx->n_pairs) {
      ConfirmArgs a{};
      a.cand = ctx->cand.as<uint2>();
      a.cinfo = ctx->cinfo.as<uint2>();
      a.block_first = ctx->block_first.as<uint32_t>();
      a.pstart = ctx->pstart.as<uint64_t>();
      a.n_cand = ctx->n_cand;
      a.n_pairs = ctx->n_pairs;
      a.tab_start = ctx->tab_start.as<uint32_t>();
      a.items = ctx->items.as<uint32_t>();
      a.pass_cnt = ctx->tab_fill.as<uint32_t>();
      a.rd_words = ctx->rd_words.as<uint64_t>();
      a.rd_x = ctx->rd_x.as<uint64_t>();
      a.len_flags = ctx->len_flags.as<uint32_t>();
      a.validmask = ctx->validmask.as<uint32_t>();
      a.tg_words = ctx->tg_words.as<uint64_t>();
      a.tg_x = ctx->tg_x.as<uint64_t>();
      a.xsum = ctx->xsum.as<uint32_t>();
      a.tg_off = ctx->tg_off.as<uint32_t>();
      a.n_targets = ctx->n_targets;
      a.nmiss = ctx->nmiss.as<int32_t>();
      a.matches = outbuf.as<uint4>();
      a.match_cap = cap;
      a.n_match = ctx->counters.as<unsigned long long>() + C_NMATCH;
      a.n_pass = ctx->counters.as<unsigned long long>() + C_NPASS;
      a.best = ctx->best.as<uint32_t>();
      a.mode = mode;
      confirm_pairs_kernel<<<grid_for(ctx->n_pairs, 256), 256, 0, ctx->stream>>>(ctx->win, a);
      LAUNCH_CHECK();
    }
    if (int rc = fetch_counters(ctx)) return rc;
    const uint64_t n = ctx->h_counters[C_NMATCH];
    if (n <= cap) {
      *n_out = n;
      return MSC_OK;
    }
    CK(outbuf.reserve(n * sizeof(uint4)));
  }
  return ctx->fail(MSC_ERR_NOMEM, "match buffer kept overflowing");
}

}  // namespace

// ===========================================================================================
extern "C" {

const char* msc_version(void) { return "muscato_b200 0.1 (sm_100a)"; }

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
    return fail("WindowWidth must be in 1..32 (wider windows are not supported by this build)");
  if (c.max_read_length < 1 || c.max_read_length > MSC_MAX_READ_LENGTH)
    return fail("MaxReadLength must be in 1..1024");
  for (int k = 0; k < c.n_windows; k++)
    if (c.windows[k] < 0) return fail("Windows: negative offset");
  if (c.match_mode != MSC_MATCH_FIRST && c.match_mode != MSC_MATCH_BEST)
    return fail("MatchMode must be 'first' or 'best'");
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

  msc_ctx* ctx = new msc_ctx();
  ctx->cfg = c;
  ctx->device = c.device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->win.nwin = c.n_windows;
  ctx->win.W = c.window_width;
  ctx->win.MRL = c.max_read_length;
  ctx->win.S = (c.max_read_length + 31) / 32;
  ctx->win.min_dinuc = c.min_dinuc;
  for (int k = 0; k < c.n_windows; k++) ctx->win.windows[k] = c.windows[k];
  bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; ok && i < 4; i++) ok = cudaEventCreate(&ctx->ev[i]) == cudaSuccess;
  ok = ok && ctx->counters.reserve(C_COUNT * sizeof(unsigned long long)) == cudaSuccess;
  ok = ok && cudaMallocHost(&ctx->h_counters, C_COUNT * sizeof(unsigned long long)) == cudaSuccess;
  // nmiss table in IEEE double exactly as cmd/muscato_confirm/main.go:198 writes it.
  std::vector<int32_t> nm(c.max_read_length + 1);
  for (int L = 0; L <= c.max_read_length; L++) {
    volatile double one_minus = 1 - c.pmatch;
    volatile double prod = one_minus * (double)L;
    nm[L] = (int32_t)prod;
  }
  ok = ok && ctx->nmiss.reserve(nm.size() * sizeof(int32_t)) == cudaSuccess;
  ok = ok && cudaMemcpy(ctx->nmiss.p, nm.data(), nm.size() * sizeof(int32_t), cudaMemcpyHostToDevice) == cudaSuccess;
  ok = ok && cudaMemset(ctx->counters.p, 0, C_COUNT * sizeof(unsigned long long)) == cudaSuccess;
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
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  DevBuf* bufs[] = {&ctx->rd_ascii, &ctx->rd_offs, &ctx->rd_words, &ctx->rd_x,  &ctx->len_flags, &ctx->validmask,
                    &ctx->tab_fp,   &ctx->tab_cnt, &ctx->tab_start, &ctx->tab_fill, &ctx->bloom, &ctx->items,
                    &ctx->tg_ascii, &ctx->tg_off,  &ctx->tg_words, &ctx->tg_x,  &ctx->xsum,      &ctx->cand,
                    &ctx->sizes,    &ctx->pstart, &ctx->cinfo, &ctx->block_first,  &ctx->match_pre, &ctx->best, &ctx->rcount,    &ctx->rstart,
                    &ctx->rfill,    &ctx->match_out, &ctx->counters, &ctx->tile_sums, &ctx->nmiss};
  for (DevBuf* b : bufs) b->release();
  if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
  for (auto& e : ctx->ev)
    if (e) cudaEventDestroy(e);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* msc_last_error(const msc_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int msc_set_reads(msc_ctx* ctx, const uint8_t* ascii, const uint64_t* offs, uint64_t n_reads) {
  if (!ctx) return MSC_ERR_STATE;
  if (n_reads && (!offs || (!ascii && offs[n_reads] != offs[0]))) return ctx->fail(MSC_ERR_INPUT, "reads: NULL buffer");
  CK(cudaSetDevice(ctx->device));
  const uint64_t nwin = (uint64_t)ctx->win.nwin;
  if (n_reads * nwin >= 0xffffffffull) return ctx->fail(MSC_ERR_INPUT, "reads: n_reads * n_windows must be < 2^32");
  uint64_t total = 0;
  if (n_reads) {
    if (offs[0] != 0) return ctx->fail(MSC_ERR_INPUT, "reads: offs[0] must be 0");
    for (uint64_t i = 0; i < n_reads; i++) {
      if (offs[i + 1] < offs[i]) return ctx->fail(MSC_ERR_INPUT, "reads: offsets not monotone at %llu", (unsigned long long)i);
      if (offs[i + 1] - offs[i] > (uint64_t)ctx->win.MRL)
        return ctx->fail(MSC_ERR_INPUT, "reads: read %llu is longer than MaxReadLength (prep_reads truncates, "
                         "cmd/muscato_prep_reads/main.go:67-69)", (unsigned long long)i);
    }
    total = offs[n_reads];
  }
  ctx->n_reads = n_reads;
  ctx->rd_ascii_bytes = total;
  const int S = ctx->win.S;
  CK(ctx->rd_ascii.reserve(total + 64));
  CK(ctx->rd_offs.reserve((n_reads + 1) * sizeof(uint64_t)));
  CK(ctx->rd_words.reserve((n_reads * S + 2) * sizeof(uint64_t)));
  CK(ctx->rd_x.reserve((n_reads * S + 2) * sizeof(uint64_t)));
  CK(ctx->len_flags.reserve((n_reads + 1) * sizeof(uint32_t)));
  CK(ctx->validmask.reserve((n_reads + 1) * sizeof(uint32_t)));
  const uint64_t kmax = std::max<uint64_t>(n_reads * nwin, 512);
  ctx->lg_slots = ceil_log2(2 * kmax);
  const int bpk = ctx->cfg.bloom_bits_per_key > 0 ? ctx->cfg.bloom_bits_per_key : 32;
  ctx->lg_bloom = std::max(10, ceil_log2((kmax * (uint64_t)bpk + 63) / 64));
  const uint64_t slots = 1ull << ctx->lg_slots;
  CK(ctx->tab_fp.reserve(slots * sizeof(uint64_t)));
  CK(ctx->tab_cnt.reserve(slots * sizeof(uint32_t)));
  CK(ctx->tab_start.reserve((slots + 1) * sizeof(uint32_t)));
  CK(ctx->tab_fill.reserve(slots * sizeof(uint32_t)));
  CK(ctx->bloom.reserve((1ull << ctx->lg_bloom) * sizeof(uint64_t)));
  CK(ctx->items.reserve((n_reads * nwin + 1) * sizeof(uint32_t)));
  CK(ctx->best.reserve((n_reads + 1) * sizeof(uint32_t)));
  // rd_words / rd_x rows are read one word past their end by extract32: keep the pad defined.
  CK(cudaMemsetAsync(ctx->rd_words.as<uint64_t>() + n_reads * S, 0, 2 * sizeof(uint64_t), ctx->stream));
  CK(cudaMemsetAsync(ctx->rd_x.as<uint64_t>() + n_reads * S, 0, 2 * sizeof(uint64_t), ctx->stream));
  if (total) CK(cudaMemcpyAsync(ctx->rd_ascii.p, ascii, total, cudaMemcpyHostToDevice, ctx->stream));
  if (n_reads) CK(cudaMemcpyAsync(ctx->rd_offs.p, offs, (n_reads + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
  else CK(cudaMemsetAsync(ctx->rd_offs.p, 0, sizeof(uint64_t), ctx->stream));
  ctx->st.h2d_bytes += total + (n_reads + 1) * sizeof(uint64_t);
  ctx->have_reads = true;
  int rc = build_reads_device(ctx);
  if (rc == MSC_OK && !ctx->cfg.keep_ascii) ctx->rd_ascii.release();
  return rc;
}

int msc_set_targets(msc_ctx* ctx, const uint8_t* ascii, const uint64_t* offs, uint64_t n_targets) {
  if (!ctx) return MSC_ERR_STATE;
  if (n_targets && (!offs || (!ascii && offs[n_targets] != offs[0]))) return ctx->fail(MSC_ERR_INPUT, "targets: NULL buffer");
  CK(cudaSetDevice(ctx->device));
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
  if (total) CK(cudaMemcpyAsync(ctx->tg_ascii.p, ascii, total, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->tg_off.p, off32.data(), (n_targets + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));  // off32 is a local
  ctx->st.h2d_bytes += total + (n_targets + 1) * sizeof(uint32_t);
  ctx->have_targets = true;
  int rc = pack_targets_device(ctx);
  if (rc == MSC_OK && !ctx->cfg.keep_ascii) ctx->tg_ascii.release();
  return rc;
}

int msc_rebuild(msc_ctx* ctx, int what) {
  if (!ctx) return MSC_ERR_STATE;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->cfg.keep_ascii) return ctx->fail(MSC_ERR_STATE, "msc_rebuild needs keep_ascii=1");
  if (what & 1) {
    if (!ctx->have_reads) return ctx->fail(MSC_ERR_STATE, "msc_rebuild: no reads set");
    if (int rc = build_reads_device(ctx)) return rc;
  }
  if (what & 2) {
    if (!ctx->have_targets) return ctx->fail(MSC_ERR_STATE, "msc_rebuild: no targets set");
    if (int rc = pack_targets_device(ctx)) return rc;
  }
  return MSC_OK;
}

int msc_screen(msc_ctx* ctx) {
  if (!ctx) return MSC_ERR_STATE;
  if (!ctx->have_reads || !ctx->have_targets) return ctx->fail(MSC_ERR_STATE, "msc_screen: set reads and targets first");
  CK(cudaSetDevice(ctx->device));
  if (ctx->cand.cap == 0) CK(ctx->cand.reserve(std::max<uint64_t>(1u << 20, ctx->n_bases / 32) * sizeof(uint2)));
  int blocks_per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, scan_targets_kernel, kScanBlock, 0));
  const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(ctx->n_tiles, (uint64_t)ctx->sm_count * std::max(1, blocks_per_sm)));
  for (int attempt = 0; attempt < 3; attempt++) {
    if (int rc = zero_counter(ctx, C_NCAND)) return rc;
    if (int rc = zero_counter(ctx, C_BLOOMPASS)) return rc;
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (ctx->n_tiles && ctx->n_reads) {
      ScanArgs a{};
      a.tg_words = ctx->tg_words.as<uint64_t>();
      a.tg_x = ctx->tg_x.as<uint64_t>();
      a.xsum = ctx->xsum.as<uint32_t>();
      a.n_bases = ctx->n_bases;
      a.n_tiles = ctx->n_tiles;
      a.bloom = ctx->bloom.as<uint2>();
      a.lg_bloom = ctx->lg_bloom;
      a.tab_fp = ctx->tab_fp.as<uint64_t>();
      a.lg_slots = ctx->lg_slots;
      a.cand = ctx->cand.as<uint2>();
      a.cand_cap = ctx->cand.cap / sizeof(uint2);
      a.n_cand = ctx->counters.as<unsigned long long>() + C_NCAND;
      a.n_bloom_pass = ctx->counters.as<unsigned long long>() + C_BLOOMPASS;
      a.W = ctx->win.W;
      scan_targets_kernel<<<grid, kScanBlock, 0, ctx->stream>>>(a);
      LAUNCH_CHECK();
    }
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    if (int rc = fetch_counters(ctx)) return rc;
    const float ms = elapsed(ctx, 0, 1);
    ctx->st.ms_scan += ms;
    ctx->st.ms_scan_kernel = ms;
    ctx->n_cand = ctx->h_counters[C_NCAND];
    if (ctx->n_cand <= ctx->cand.cap / sizeof(uint2)) {
      ctx->st.n_candidates = ctx->n_cand;
      ctx->st.positions_probed = ctx->n_bases;
      ctx->st.reserved_f[0] = (float)ctx->h_counters[C_BLOOMPASS];
      ctx->have_cand = true;
      ctx->have_confirm = ctx->have_combine = false;
      return MSC_OK;
    }
    CK(ctx->cand.reserve(ctx->n_cand * sizeof(uint2)));
  }
  return ctx->fail(MSC_ERR_NOMEM, "candidate buffer kept overflowing");
}

static int expand_candidates(msc_ctx* ctx) {
  CK(ctx->sizes.reserve((ctx->n_cand + 1) * sizeof(uint32_t)));
  CK(ctx->cinfo.reserve((ctx->n_cand + 1) * sizeof(uint2)));
  CK(ctx->pstart.reserve((ctx->n_cand + 2) * sizeof(uint64_t)));
  if (ctx->n_cand) {
    cand_prepare_kernel<<<grid_for(ctx->n_cand, 256), 256, 0, ctx->stream>>>(
        ctx->cand.as<uint2>(), ctx->n_cand, ctx->tab_cnt.as<uint32_t>(), ctx->tg_off.as<uint32_t>(), ctx->n_targets,
        ctx->win.W, ctx->cinfo.as<uint2>(), ctx->sizes.as<uint32_t>());
    LAUNCH_CHECK();
  }
  if (int rc = device_exclusive_scan<uint64_t>(ctx, ctx->sizes.as<uint32_t>(), ctx->n_cand, ctx->pstart.as<uint64_t>(), true))
    return rc;
  if (int rc = fetch_counters(ctx)) return rc;
  ctx->n_pairs = ctx->h_counters[C_SCANTOTAL];
  ctx->st.n_pairs = ctx->n_pairs;
  const uint64_t n_blocks = (ctx->n_pairs + 255) / 256;
  CK(ctx->block_first.reserve((n_blocks + 2) * sizeof(uint32_t)));
  if (ctx->n_pairs) {
    pair_block_starts_kernel<<<grid_for(n_blocks + 1, 256), 256, 0, ctx->stream>>>(
        ctx->pstart.as<uint64_t>(), ctx->n_cand, ctx->n_pairs, n_blocks, ctx->block_first.as<uint32_t>());
    LAUNCH_CHECK();
  }
  return MSC_OK;
}

int msc_confirm(msc_ctx* ctx) {
  if (!ctx) return MSC_ERR_STATE;
  if (!ctx->have_cand) return ctx->fail(MSC_ERR_STATE, "msc_confirm: run msc_screen first");
  CK(cudaSetDevice(ctx->device));
  CK(cudaEventRecord(ctx->ev[0], ctx->stream));
  if (int rc = expand_candidates(ctx)) return rc;
  CK(cudaEventRecord(ctx->ev[1], ctx->stream));
  if (ctx->match_pre.cap == 0) CK(ctx->match_pre.reserve((size_t)(1u << 20) * sizeof(uint4)));
  if (int rc = run_confirm_kernel(ctx, 0, ctx->match_pre, &ctx->n_match_pre)) return rc;
  CK(cudaEventRecord(ctx->ev[2], ctx->stream));
  // MaxMatches pre-check (cmd/muscato_confirm/main.go:233-242, :424-448): truncation can only
  // happen in a key group with more than MaxMatches passing pairs.
  const uint64_t n_pass = ctx->h_counters[C_NPASS];
  uint64_t n_over = 0;
  if (n_pass > (uint64_t)ctx->cfg.max_matches) {
    if (int rc = zero_counter(ctx, C_NOVER)) return rc;
    const uint64_t slots = 1ull << ctx->lg_slots;
    overflow_count_kernel<<<grid_for(slots, 256), 256, 0, ctx->stream>>>(
        ctx->tab_fill.as<uint32_t>(), slots, (unsigned long long)ctx->cfg.max_matches,
        ctx->counters.as<unsigned long long>() + C_NOVER);
    LAUNCH_CHECK();
    if (int rc = fetch_counters(ctx)) return rc;
    n_over = ctx->h_counters[C_NOVER];
  }
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->st.ms_expand += elapsed(ctx, 0, 1);
  ctx->st.ms_confirm += elapsed(ctx, 1, 2);
  ctx->st.n_pass = n_pass;
  ctx->st.n_matches_pre = ctx->n_match_pre;
  ctx->st.n_overflow_groups = n_over;
  if (n_over)
    return ctx->fail(MSC_ERR_CONFIG,
                     "%llu key group(s) exceed MaxMatches=%lld passing pairs; the order-dependent truncation of "
                     "cmd/muscato_confirm/main.go:424-448 is not implemented yet -- raise MaxMatches",
                     (unsigned long long)n_over, (long long)ctx->cfg.max_matches);
  ctx->have_confirm = true;
  ctx->have_combine = false;
  return MSC_OK;
}

void* msc_best_device(msc_ctx* ctx) { return (ctx && ctx->have_confirm) ? ctx->best.p : nullptr; }

void* msc_matches_device(msc_ctx* ctx, uint64_t* n) {
  if (!ctx || !ctx->have_combine) return nullptr;
  if (n) *n = ctx->n_match;
  return ctx->match_out.p;
}

int msc_combine(msc_ctx* ctx) {
  if (!ctx) return MSC_ERR_STATE;
  if (!ctx->have_confirm) return ctx->fail(MSC_ERR_STATE, "msc_combine: run msc_confirm first");
  CK(cudaSetDevice(ctx->device));
  const uint64_t U = ctx->n_reads, n = ctx->n_match_pre;
  if (n >= 0xffffffffull) return ctx->fail(MSC_ERR_NOMEM, "more than 2^32 matches in one batch");
  CK(ctx->rcount.reserve((U + 1) * sizeof(uint32_t)));
  CK(ctx->rstart.reserve((U + 2) * sizeof(uint32_t)));
  CK(ctx->rfill.reserve((U + 1) * sizeof(uint32_t)));
  CK(ctx->match_out.reserve((n + 1) * sizeof(uint4)));
  CK(cudaEventRecord(ctx->ev[0], ctx->stream));
  CK(cudaMemsetAsync(ctx->rcount.p, 0, (U + 1) * sizeof(uint32_t), ctx->stream));
  CK(cudaMemsetAsync(ctx->rfill.p, 0, (U + 1) * sizeof(uint32_t), ctx->stream));
  if (n) {
    combine_count_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(ctx->match_pre.as<uint4>(), n, ctx->best.as<uint32_t>(),
                                                                    (uint32_t)ctx->cfg.mmtol, ctx->rcount.as<uint32_t>());
    LAUNCH_CHECK();
  }
  if (int rc = device_exclusive_scan<uint32_t>(ctx, ctx->rcount.as<uint32_t>(), U, ctx->rstart.as<uint32_t>(), true)) return rc;
  if (n) {
    combine_scatter_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(ctx->match_pre.as<uint4>(), n, ctx->best.as<uint32_t>(),
                                                                      (uint32_t)ctx->cfg.mmtol, ctx->rstart.as<uint32_t>(),
                                                                      ctx->rfill.as<uint32_t>(), ctx->match_out.as<uint4>());
    LAUNCH_CHECK();
  }
  CK(cudaEventRecord(ctx->ev[1], ctx->stream));
  if (int rc = fetch_counters(ctx)) return rc;
  ctx->n_match = ctx->h_counters[C_SCANTOTAL];
  ctx->st.n_matches = ctx->n_match;
  ctx->st.ms_combine += elapsed(ctx, 0, 1);
  ctx->have_combine = true;
  return MSC_OK;
}

int msc_fetch_matches(msc_ctx* ctx, msc_match** out, uint64_t* n) {
  if (!ctx || !out || !n) return MSC_ERR_STATE;
  if (!ctx->have_combine) return ctx->fail(MSC_ERR_STATE, "msc_fetch_matches: run msc_combine first");
  CK(cudaSetDevice(ctx->device));
  *n = ctx->n_match;
  msc_match* h = (msc_match*)malloc(std::max<uint64_t>(1, ctx->n_match) * sizeof(msc_match));
  if (!h) return ctx->fail(MSC_ERR_NOMEM, "host allocation failed");
  if (ctx->n_match) {
    static_assert(sizeof(msc_match) == sizeof(uint4), "msc_match layout");
    cudaError_t e = cudaMemcpyAsync(h, ctx->match_out.p, ctx->n_match * sizeof(msc_match), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      free(h);
      return ctx->fail(MSC_ERR_CUDA, "D2H of matches failed: %s", cudaGetErrorString(e));
    }
    ctx->st.d2h_bytes += ctx->n_match * sizeof(msc_match);
    // The device groups by read (counting sort on read_id); order inside a read group is made
    // deterministic here.  Groups are short, so this is a linear pass in practice.
    auto by_gene_pos = [](const msc_match& a, const msc_match& b) {
      return a.gene_id != b.gene_id ? a.gene_id < b.gene_id : a.pos < b.pos;
    };
    for (uint64_t i = 0; i < ctx->n_match;) {
      uint64_t j = i + 1;
      while (j < ctx->n_match && h[j].read_id == h[i].read_id) j++;
      if (j - i > 1) std::sort(h + i, h + j, by_gene_pos);
      i = j;
    }
  }
  *out = h;
  return MSC_OK;
}

int msc_run(msc_ctx* ctx) {
  if (int rc = msc_screen(ctx)) return rc;
  if (int rc = msc_confirm(ctx)) return rc;
  return msc_combine(ctx);
}

int msc_get_stats(const msc_ctx* ctx, msc_stats* out) {
  if (!ctx || !out) return MSC_ERR_STATE;
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
  std::vector<uint32_t> items(ctx->n_keys);
  if (ctx->n_keys) CK(cudaMemcpy(items.data(), ctx->items.p, ctx->n_keys * sizeof(uint32_t), cudaMemcpyDeviceToHost));
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
  if (int rc = expand_candidates(ctx)) return rc;
  DevBuf tmp;
  CK(tmp.reserve((size_t)(1u << 16) * sizeof(uint4)));
  uint64_t cnt = 0;
  int rc = run_confirm_kernel(ctx, 1, tmp, &cnt);
  ctx->have_confirm = ctx->have_combine = false;  // the pair kernel scratch (best / pass counts) was reused
  if (rc) { tmp.release(); return rc; }
  msc_cand_rec* h = (msc_cand_rec*)malloc(std::max<uint64_t>(1, cnt) * sizeof(msc_cand_rec));
  if (!h) { tmp.release(); return ctx->fail(MSC_ERR_NOMEM, "host allocation failed"); }
  static_assert(sizeof(msc_cand_rec) == sizeof(uint4), "msc_cand_rec layout");
  cudaError_t e = cnt ? cudaMemcpy(h, tmp.p, cnt * sizeof(uint4), cudaMemcpyDeviceToHost) : cudaSuccess;
  tmp.release();
  if (e != cudaSuccess) { free(h); return ctx->fail(MSC_ERR_CUDA, "D2H failed: %s", cudaGetErrorString(e)); }
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
