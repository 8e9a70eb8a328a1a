// scan.cuh -- kernel (2): streaming scan of the 2-bit target database.
//
// Replaces search/processSeq/checkWin/harvest (cmd/muscato_screen/main.go:220-480): the
// reference rolls NumHash buzhash32 hashes over every target base and probes one
// BloomSize-bit array per window; here every target position's W-mer IS its key (W<=32),
// obtained by a funnel shift of two adjacent packed words, and ONE probe of a blocked Bloom
// word (8 bytes) per position covers all windows.  Positions that pass go to the exact key
// table; hits are compacted as (table slot, global target position) with warp-aggregated
// atomics.  Window/target-boundary and position-0 rules (processSeq :294-365) are applied
// by the confirm kernel, which knows the gene of each candidate.
//
// Data movement: every WARP owns a stream of 32-word tiles (1024 bases + halo) that the TMA
// engine stages into the warp's shared-memory slice (1-D cp.async.bulk + one mbarrier per
// buffer, double buffered); warps never wait for each other -- there is no block-wide barrier
// in the loop, so a warp that is draining hits (latency bound) overlaps with warps that are
// probing (issue bound).  The Bloom/table probes are scattered 8-byte reads served from L2
// when the filter fits (it is sized 32-64 bits per key) and from HBM sectors otherwise.
#pragma once
#include "common.cuh"

namespace msc {

constexpr int kScanBlock = 256;
constexpr int kScanWarps = kScanBlock / 32;
constexpr int kTileWords = 256;                       // granularity of the target buffers (8 warp tiles)
constexpr int kWarpTileWords = 32;                    // one 32-base word per lane
constexpr int kWarpCopyWords = kWarpTileWords + 2;    // halo word + 1 (byte count multiple of 16)
constexpr int kWarpSmemWords = kWarpTileWords + 8;    // keeps every buffer 64-byte aligned
constexpr int kStageCap = 128;                        // per-warp candidate staging (entries)
constexpr int kGeneBlockShift = 10;                   // granularity of the position -> target index (msc_set_targets)
constexpr int kGeneTab = 64;                          // target offsets a warp keeps in shared memory per tile
#ifndef MSC_SCAN_BATCH
#define MSC_SCAN_BATCH 8
#endif
#ifndef MSC_SCAN_CTAS
#define MSC_SCAN_CTAS 4
#endif
constexpr int kProbeBatch = MSC_SCAN_BATCH;           // Bloom probes in flight per lane

struct ScanArgs {
  const uint64_t* tg_words;
  const uint64_t* tg_x;
  const uint32_t* xsum;
  const unsigned long long* targets_have_x;  // device flag: 0 = no target word contains X (no summary look-ups at all)
  uint64_t n_bases;
  uint64_t n_tiles;       // 256-word tiles; the buffers are padded to n_tiles * 256 + 64 words
  const uint2* bloom;
  BloomGeom geom;
  uint32_t mul[8];        // mul[j] = 1 << (32 - 2m - 2j): m-mer j of a key to the top of a word
  const uint8_t* tab;     // key table: 128-byte buckets (common.cuh)
  uint64_t n_buckets;
  // the candidate's 32-byte record for the pair kernel (two uint4, see flush_stage) and its key group's size
  uint4* cinfo;
  uint32_t* sizes;
  const uint32_t* tg_off;    // n_targets + 1 base offsets of the concatenated stream
  const uint32_t* blk2gene;  // target that holds base b << kGeneBlockShift
  uint64_t n_targets;
  unsigned long long cand_cap;
  unsigned long long* n_cand;
  unsigned long long* n_bloom_pass;
  unsigned long long* n_dummy;  // exact front: candidate slots that were reserved and stayed empty (size 0)
  int W;
  int alu_masks;          // 1 = Bloom bit masks by arithmetic instead of the shared-memory pattern table (MSC_SCAN_ALU_MASKS)
  int prefetch;           // 1 = request every queued position's bucket line into the L2 before the drain (MSC_SCAN_PREFETCH)
  int pass;               // exact front (geom.direct): the slice of the bitmap this launch tests, 0 .. 2^lg_pass - 1
  int stream_tab;         // 1 = table beyond the L2: bucket lines and candidate records carry the evict-first L2 policy
};

// Shared memory of one scan CTA (dynamic: 50 KB, above the 48 KB static limit).
struct ScanSmem {
  alignas(128) uint64_t tiles[kScanWarps][2][kWarpSmemWords];
  alignas(16) uint4 stage_rec[kScanWarps][kStageCap];  // the key group record of every staged candidate
  alignas(8) uint64_t bars[kScanWarps][2];
  uint2 stage[kScanWarps][kStageCap];  // found (slot, position) pairs, flushed when nearly full
  uint32_t stage_g[kScanWarps][kStageCap];  // the target that holds each staged position
  uint16_t queue[kScanWarps][1024];
  uint32_t gtab[kScanWarps][kGeneTab + 4];  // tg_off[g0 .. g0 + kGeneTab] of the current tile (g0 = target of its first block)
  uint32_t pattern[1024];              // bloom_pattern(): the two low-half bits of a key
};

// Extract the W-mer that starts at base j (0..31) of the word pair (lo, hi).
__device__ __forceinline__ uint64_t window_at(uint64_t lo, uint64_t hi, unsigned j, uint64_t kmask) {
  return (j == 0 ? lo : ((lo >> (2 * j)) | (hi << (64 - 2 * j)))) & kmask;
}

// Two phases per warp tile, so that the scattered probes never serialise behind divergent hit
// handling:
//   phase 1  the warp walks the 32 words of its tile; in every step the 32 lanes test the 32
//            CONSECUTIVE positions of one word against the Bloom front (lane = position), so
//            that lanes whose W-mers share a minimiser hit the same 32-byte sector and coalesce
//            into one L1 wavefront / one L2 request (common.cuh, "locality aware").  kProbeBatch steps
//            are in flight per lane.  The ballot of step i is kept by lane i: after the walk
//            lane i owns the 32-bit pass mask of word i;
//   phase 2  the warp compacts its passes into a shared-memory queue and drains it 32 at a time:
//            lane i re-derives the key of queued position i with two shuffles, looks it up in the
//            exact table, and the warp appends the found (slot, position) pairs with ONE atomic.
// KW: 0 = W <= 16 (the key arithmetic of phase 1 is 32 bit), 1 = W <= 32, 2 = wide window
// (32 < W <= 64, two key words).  WN: number of competing m-mers.  (The exact front of BloomGeom::direct has its own
// kernel, scan_direct.cuh.)
template <int KW, int WN>
__global__ void __launch_bounds__(kScanBlock, MSC_SCAN_CTAS) scan_targets_kernel(const ScanArgs a) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char scan_smem[];
  ScanSmem& sm = *reinterpret_cast<ScanSmem*>(scan_smem);
  auto& tiles = sm.tiles;
  auto& bars = sm.bars;
  auto& queue = sm.queue;
  auto& stage = sm.stage;
  auto& stage_rec = sm.stage_rec;
  auto& pattern = sm.pattern;
  uint32_t* gt = sm.gtab[threadIdx.x >> 5];
  for (int i = threadIdx.x; i < 1024; i += kScanBlock) pattern[i] = bloom_pattern((uint32_t)i);
  __syncthreads();
  const int tid = threadIdx.x;
  const unsigned lane = tid & 31u, warp = tid >> 5;
  constexpr uint32_t kBytes = kWarpCopyWords * sizeof(uint64_t);
  uint64_t* bar = bars[warp];
  if (lane == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  __syncwarp();

  const bool any_x = *a.targets_have_x != 0ull;
  constexpr bool K32 = KW == 0;
  constexpr bool WIDE = KW == 2;
  const uint64_t pol_once = l2_policy_evict_first();
  const bool stream_tab = a.stream_tab != 0;
  const uint64_t kmask = low_bases_mask(min(a.W, 32));
  const uint64_t kmask1 = WIDE ? low_bases_mask(a.W - 32) : 0ull;  // bases 32..W-1 of a wide window
  const uint32_t xr = a.geom.xr;
  const int gm = a.geom.m;
  const int lg_words = a.geom.lg_words;
  const int lg_blk = a.geom.lg_blk;
  // Static, balanced split: the stream is cut into units of kProbeBatch words and every warp of
  // the grid gets a contiguous run of units whose length differs by at most one between warps;
  // the warp walks its run in tiles of up to 32 words (the last one may be shorter).
  const uint64_t n_words = (a.n_bases + 31) >> 5;
  const uint64_t n_units = (n_words + kProbeBatch - 1) / kProbeBatch;
  const uint64_t n_warps = (uint64_t)kScanWarps * gridDim.x, gw = (uint64_t)blockIdx.x * kScanWarps + warp;
  uint64_t w0 = (gw * n_units / n_warps) * kProbeBatch;                                  // first word of the current tile
  const uint64_t w_end = ((gw + 1) * n_units / n_warps) * kProbeBatch;  // end of the run (buffers are padded)
  if (lane == 0 && w0 < w_end) {
    mbar_arrive_expect_tx(&bar[0], kBytes);
    bulk_copy_g2s(tiles[warp][0], a.tg_words + w0, kBytes, &bar[0]);
  }
  uint32_t phases = 0;
  int buf = 0;
  uint32_t n_pass = 0;
  uint16_t* q16 = queue[warp];
  uint2* st = stage[warp];
  uint32_t n_st = 0;  // staged candidates of this warp (warp-uniform)
  // One global atomic per flush instead of one per drain round: same-address atomics serialise.
  uint4* st_rec = stage_rec[warp];
  uint32_t* st_g = sm.stage_g[warp];
  // Flush: one global atomic reserves the run; every staged candidate is written as the pair kernel wants it --
  // (slot, position), its key group's size, and ONE 32-byte record (two uint4):
  //   (global position of the window, window start p inside its target, global end of the target, read record of item 0)
  //   (item 0 of the key group, CSR start of the further items, target index, table slot)
  // so the expansion needs no pass over the candidates of its own.  The target of a position comes from the tile's
  // slice of the offset table in shared memory (gt[]: loaded once per tile; the stage is flushed at the end of every
  // tile, so all staged positions belong to the current one) -- per-candidate look-ups in global memory were measured
  // to miss the L2 this kernel's random traffic keeps flushing.  A tile with more than kGeneTab target starts falls
  // back to the global search.  A W-mer that straddles a target boundary is not a window of any target (processSeq
  // only rolls inside one target, cmd/muscato_screen/main.go:319): size 0.
  uint32_t g0 = 0;     // target that holds the first base of the tile's first 2^kGeneBlockShift block
  bool gt_ok = false;  // gt[] covers the whole tile
  uint32_t gb1 = 0, gb2 = 0, gb3 = 0;  // tg_off[g0 + 1 .. g0 + 3] (warp-uniform copies of gt[1..3])
  auto flush_stage = [&]() {
    unsigned long long out0 = 0;
    if (lane == 0) out0 = atomicAdd(a.n_cand, (unsigned long long)n_st);
    out0 = __shfl_sync(0xffffffffu, out0, 0);
#pragma unroll 2
    for (uint32_t i = lane; i < n_st; i += 32) {
      const uint2 e = st[i];
      const uint4 rec = st_rec[i];
      // the target was resolved when the candidate was staged (target_of below: the tile's slice of the offset table);
      // its two offsets come through the L1 (neighbouring candidates share them)
      const uint32_t g = st_g[i];
      const uint32_t goff = __ldg(a.tg_off + g), gend = __ldg(a.tg_off + g + 1);
      const unsigned long long o = out0 + i;
      if (o < a.cand_cap) {
        const uint32_t sz = ((uint64_t)e.y + (uint64_t)a.W <= (uint64_t)gend) ? 1u + rec.w : 0u;
        if (stream_tab) {
          // written once, read once by the expansion long after the L2 has turned over: first out
          stg128_hint(a.cinfo + 2 * o, make_uint4(e.y, e.y - goff, gend, rec.y), pol_once);
          stg128_hint(a.cinfo + 2 * o + 1, make_uint4(rec.x, rec.z, g, e.x), pol_once);
          stg32_hint(a.sizes + o, sz, pol_once);
        } else {
          a.cinfo[2 * o] = make_uint4(e.y, e.y - goff, gend, rec.y);
          a.cinfo[2 * o + 1] = make_uint4(rec.x, rec.z, g, e.x);
          a.sizes[o] = sz;
        }
      }
    }
    __syncwarp();
    n_st = 0;
  };
  // Target that holds position `pos` of the current tile: the tile's first three targets from registers, its first
  // kGeneTab from the warp's shared-memory slice of the offset table, anything beyond by a search in global memory.
  auto target_of = [&](uint32_t pos) -> uint32_t {
    if (pos < gb3) return g0 + (pos >= gb1) + (pos >= gb2);
    if (gt_ok) {
      // number of offsets gt[1..kGeneTab] <= position (ascending): branch-free binary search in shared memory
      uint32_t c = 0;
#pragma unroll
      for (int step = kGeneTab / 2; step >= 1; step >>= 1) c += (gt[c + step] <= pos) ? step : 0;
      return g0 + c;
    }
    const uint64_t g_lo = __ldg(a.blk2gene + (pos >> kGeneBlockShift)), g_hi = __ldg(a.blk2gene + (pos >> kGeneBlockShift) + 1);
    return (uint32_t)(upper_bound_dev<uint32_t>(a.tg_off, g_lo + 1, g_hi + 1, pos) - 1);
  };
  uint32_t g0_next = w0 < w_end ? __ldg(a.blk2gene + ((w0 * 32ull) >> kGeneBlockShift)) : 0u;
  for (; w0 < w_end; w0 += kWarpTileWords) {
    const uint64_t wn = w0 + kWarpTileWords;
    if (lane == 0 && wn < w_end) {
      mbar_arrive_expect_tx(&bar[buf ^ 1], kBytes);
      bulk_copy_g2s(tiles[warp][buf ^ 1], a.tg_words + wn, kBytes, &bar[buf ^ 1]);
    }
    const int tile_words = (int)min((uint64_t)kWarpTileWords, w_end - w0);  // multiple of kProbeBatch
    // the tile's slice of the target offsets: requested now, consumed only when phase 2 starts (the loads overlap the
    // tile's arrival and the whole of phase 1)
    g0 = g0_next;
    if (wn < w_end) g0_next = __ldg(a.blk2gene + ((wn * 32ull) >> kGeneBlockShift));  // for the next tile: no dependent wait
    const uint32_t gt_o0 = __ldg(a.tg_off + min((uint64_t)g0 + lane, a.n_targets));
    const uint32_t gt_o1 = __ldg(a.tg_off + min((uint64_t)g0 + 32u + lane, a.n_targets));
    const uint32_t gt_o2 = __ldg(a.tg_off + min((uint64_t)g0 + 64u, a.n_targets));
    mbar_wait(&bar[buf], (phases >> buf) & 1u);
    phases ^= 1u << buf;
    const uint64_t* tile = tiles[warp][buf];

    // lane i owns word i of the tile: its pass mask and its X flag
    uint32_t mask = 0;
    // X summary bits of word w and w+1 (xsum is padded): bit i of xwords = word i of the tile needs the X path
    unsigned xwords = 0;
    if (any_x) {
      const uint64_t w = w0 + (uint64_t)lane;
      const uint32_t xs0 = __ldg(a.xsum + (w >> 5));
      const uint32_t xs1 = __ldg(a.xsum + ((w + 1) >> 5));
      uint32_t anyx = (xs0 >> (unsigned)(w & 31u)) | (xs1 >> (unsigned)((w + 1) & 31u));
      if (WIDE) anyx |= __ldg(a.xsum + ((w + 2) >> 5)) >> (unsigned)((w + 2) & 31u);  // a wide window reaches word w + 2
      xwords = __ballot_sync(0xffffffffu, anyx & 1u);
    }
    if (tile_words < 32) xwords &= (1u << tile_words) - 1u;
    const unsigned xwords_all = xwords;
    {
      const uint32_t* t32 = reinterpret_cast<const uint32_t*>(tile) + (lane >> 4);
      const unsigned sh = (2u * lane) & 31u;
#pragma unroll 1
      for (int ib = 0; ib < tile_words; ib += kProbeBatch) {
        uint32_t h[kProbeBatch];
        uint2 bw[kProbeBatch];
#pragma unroll
        for (int i = 0; i < kProbeBatch; i++) {
          const int wi = ib + i;  // word of the tile; this lane tests position `lane` of it
          uint32_t prex;
          if (K32) {
            prex = (__funnelshift_r(t32[2 * wi], t32[2 * wi + 1], sh) & (uint32_t)kmask) ^ xr;
            h[i] = bloom_hash32<true>(prex, 0u);
          } else {
            const uint64_t key = window_at(tile[wi], tile[wi + 1], lane, kmask);
            prex = (uint32_t)key ^ xr;
            if (WIDE) h[i] = bloom_hash32<false>(prex, wide_khi(key, window_at(tile[wi + 1], tile[wi + 2], lane, kmask1)));
            else h[i] = bloom_hash32<false>(prex, (uint32_t)(key >> 32));
          }
          const uint32_t sec = bloom_sector_of(bloom_min_mmer<WN>(prex, a.mul), gm, lg_words, lg_blk);
          bw[i] = __ldg(a.bloom + __funnelshift_l(h[i], sec, lg_blk));  // (block << lg_blk) | (h >> (32 - lg_blk))
        }
#pragma unroll
        for (int i = 0; i < kProbeBatch; i++) {
          // same masks as bloom_masks32(): pattern table look-up + rotate
          // pattern table in shared memory (L2-resident regime: the ALU pipe is the limit) or two shifts (HBM regime:
          // the shared-memory pipe's latency is on the warp's critical path, the ALU is idle)
          const uint32_t mlo = a.alu_masks ? bloom_pattern((h[i] >> 2) & 1023u)
                                           : *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(pattern) + (h[i] & 0xFFCu));
          const uint32_t mhi = __funnelshift_l(mlo, mlo, h[i] >> 12);
          const unsigned b = __ballot_sync(0xffffffffu, ((~bw[i].x & mlo) | (~bw[i].y & mhi)) == 0u);
          if ((int)lane == ib + i) mask = b;
        }
      }
      // Words that contain X (or whose successor does): redo them with the X mask folded in.
      while (xwords) {
        const int wi = __ffs(xwords) - 1;
        xwords &= xwords - 1;
        const uint64_t xl = __ldg(a.tg_x + w0 + wi), xh = __ldg(a.tg_x + w0 + wi + 1);
        const uint64_t key = window_at(tile[wi], tile[wi + 1], lane, kmask), xm0 = window_at(xl, xh, lane, kmask);
        uint64_t key1 = 0, xm1 = 0;
        if (WIDE) {
          key1 = window_at(tile[wi + 1], tile[wi + 2], lane, kmask1);
          xm1 = window_at(xh, __ldg(a.tg_x + w0 + wi + 2), lane, kmask1);
        }
        const uint64_t xm = xm0 | xm1;
        uint64_t widx;
        uint32_t mlo, mhi;
        bloom_locate(key, xm, xm ? (WIDE ? key_fp_wide(key, key1, xm0, xm1) : key_fp(key, xm0)) : 0ull, a.W, a.geom, widx,
                     mlo, mhi, key1);
        const uint2 bwx = __ldg(a.bloom + widx);
        const unsigned b = __ballot_sync(0xffffffffu, ((bwx.x & mlo) == mlo) & ((bwx.y & mhi) == mhi));
        if ((int)lane == wi) mask = b;
      }
      const uint64_t gbase = (w0 + (uint64_t)lane) * 32ull;
      if ((int)lane >= tile_words || gbase >= a.n_bases) mask = 0;
      else if (a.n_bases - gbase < 32) mask &= (1u << (unsigned)(a.n_bases - gbase)) - 1u;
    }

    // ---- phase 2: warp-cooperative drain of the Bloom passes -------------------------------
    const uint32_t cnt = __popc(mask);
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if ((int)lane >= o) incl += v;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (total) {
      gt[lane] = gt_o0;
      gt[32 + lane] = gt_o1;
      if (lane == 0) gt[kGeneTab] = gt_o2;
      // covered when the offset after the table lies beyond the tile's last base (offsets past the last target repeat
      // the total, which is beyond every position)
      gt_ok = (uint64_t)gt_o2 > w0 * 32ull + (uint64_t)(32 * kWarpTileWords - 1);
      gb1 = __shfl_sync(0xffffffffu, gt_o0, 1);
      gb2 = __shfl_sync(0xffffffffu, gt_o0, 2);
      gb3 = __shfl_sync(0xffffffffu, gt_o0, 3);
      n_pass += cnt;
      uint32_t at = incl - cnt;
      uint32_t m = mask;
      while (m) {
        const uint32_t j = __ffs(m) - 1;
        m &= m - 1;
        q16[at++] = (uint16_t)((lane << 5) | j);
      }
      __syncwarp();
      const uint64_t wbase = w0 * 32ull;  // first position of this warp tile
      // fingerprint of a queued position (e = word << 5 | base), as the table stores it
      auto entry_fp = [&](uint32_t e) -> uint64_t {
        const unsigned src = e >> 5, j = e & 31u;
        uint64_t xm = 0, xm1 = 0;
        const bool hasx = (xwords_all >> src) & 1u;
        if (hasx) xm = window_at(__ldg(a.tg_x + w0 + src), __ldg(a.tg_x + w0 + src + 1), j, kmask);
        if (WIDE) {
          if (hasx) xm1 = window_at(__ldg(a.tg_x + w0 + src + 1), __ldg(a.tg_x + w0 + src + 2), j, kmask1);
          return key_fp_wide(window_at(tile[src], tile[src + 1], j, kmask), window_at(tile[src + 1], tile[src + 2], j, kmask1), xm, xm1);
        }
        return key_fp(window_at(tile[src], tile[src + 1], j, kmask), xm);
      };
      // Pass A: the home bucket LINE of every queued position is requested into the L2 up front (prefetch: no
      // register, no scoreboard) -- all of the tile's look-ups are in flight together, and the drain below, which
      // can only keep kDrain per lane in registers, finds its lines in the L2 instead of waiting for HBM round by round.
      if (a.prefetch)
        for (uint32_t idx = lane; idx < total; idx += 32)
          prefetch_l2(bucket_ptr(a.tab, table_home_bucket(entry_fp(q16[idx]), a.n_buckets)));
      // Pass B: kDrain look-ups per lane and round: the home bucket (first four fingerprints, one 256-bit load) of
      // each is fetched before any is resolved; then the group records of the hits (same line) are fetched together.
      constexpr int kDrain = 2;
      for (uint32_t base = 0; base < total; base += 32 * kDrain) {
        uint64_t fp[kDrain], bk[kDrain], q[kDrain][4];
        uint32_t e[kDrain];
#pragma unroll
        for (int u = 0; u < kDrain; u++) {
          const uint32_t idx = base + 32 * u + lane;
          e[u] = idx < total ? q16[idx] : 0xffffffffu;
          fp[u] = 0;
          bk[u] = 0;
          q[u][0] = q[u][1] = q[u][2] = q[u][3] = 0;
          if (e[u] != 0xffffffffu) {
            fp[u] = entry_fp(e[u]);
            bk[u] = table_home_bucket(fp[u], a.n_buckets);
            ldg256(bucket_ptr(a.tab, bk[u]), q[u][0], q[u][1], q[u][2], q[u][3]);  // the first four fingerprints of the home bucket
          }
        }
        int r[kDrain];
#pragma unroll
        for (int u = 0; u < kDrain; u++) {
          // slots fill in order, so the fifth fingerprint only matters when the first four are taken
          // by other keys (rare at the table's load factor): it is fetched on demand
          r[u] = 5;  // 0..4 found, 5 = not in the table
          if (fp[u]) {
            while (true) {
              if (q[u][0] == fp[u]) { r[u] = 0; break; }
              if (q[u][1] == fp[u]) { r[u] = 1; break; }
              if (q[u][2] == fp[u]) { r[u] = 2; break; }
              if (q[u][3] == fp[u]) { r[u] = 3; break; }
              if ((q[u][0] == 0ull) | (q[u][1] == 0ull) | (q[u][2] == 0ull) | (q[u][3] == 0ull)) break;
              const uint64_t q4 = __ldg(reinterpret_cast<const unsigned long long*>(bucket_ptr(a.tab, bk[u]) + 32));
              if (q4 == fp[u]) { r[u] = 4; break; }
              if (q4 == 0ull) break;
              bk[u] = bk[u] + 1 == a.n_buckets ? 0 : bk[u] + 1;  // bucket full of other keys: walk on
              ldg256(bucket_ptr(a.tab, bk[u]), q[u][0], q[u][1], q[u][2], q[u][3]);
            }
          }
        }
        // the group records of the hits sit in the bucket lines the look-ups have just brought in: fetched NOW, while
        // the lines are in the L2 for certain (fetched at flush time they were measured to come from HBM a second time)
        uint4 rr[kDrain];
#pragma unroll
        for (int u = 0; u < kDrain; u++) {
          rr[u] = make_uint4(0u, 0u, 0u, 0u);
          if (r[u] < kBucketSlots) {
            const uint4* rp = reinterpret_cast<const uint4*>(bucket_ptr(a.tab, bk[u]) + kBucketRecOff) + r[u];
            rr[u] = stream_tab ? ldg128_last_use(rp, pol_once) : __ldg(rp);  // the line's last use: first out of the L2
          }
        }
#pragma unroll
        for (int u = 0; u < kDrain; u++) {
          if (base + 32 * u >= total) break;  // warp-uniform
          const bool hit = r[u] < kBucketSlots;
          const unsigned found = __ballot_sync(0xffffffffu, hit);
          if (found) {
            if (hit) {
              const uint32_t at = n_st + __popc(found & ((1u << lane) - 1u));
              st[at] = make_uint2((uint32_t)(bk[u] * kBucketSlots + (uint64_t)r[u]), (uint32_t)(wbase + e[u]));
              st_rec[at] = rr[u];
              st_g[at] = target_of((uint32_t)(wbase + e[u]));
            }
            n_st += __popc(found);
            __syncwarp();
            if (n_st > kStageCap - 32) flush_stage();
          }
        }
      }
      __syncwarp();  // (the stage outlives the tile: every entry carries its target)
    }
    __syncwarp();  // all lanes are done with tile[buf] before lane 0 lets the TMA engine refill it
    buf ^= 1;
  }
  if (n_st) flush_stage();
  n_pass = __reduce_add_sync(0xffffffffu, n_pass);
  if (lane == 0 && n_pass) atomicAdd(a.n_bloom_pass, (unsigned long long)n_pass);
}

}  // namespace msc
