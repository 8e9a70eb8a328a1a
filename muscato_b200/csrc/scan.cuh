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
  uint64_t n_bases;
  uint64_t n_tiles;       // 256-word tiles; the buffers are padded to n_tiles * 256 + 64 words
  const uint2* bloom;
  BloomGeom geom;
  uint32_t mul[8];        // mul[j] = 1 << (32 - 2m - 2j): m-mer j of a key to the top of a word
  const uint8_t* tab;     // key table: 128-byte buckets (common.cuh)
  uint64_t n_buckets;
  uint2* cand;            // (slot, global target position)
  uint4* cmeta;           // the slot's record {item0, rmx0, start, cnt}, same index as cand
  unsigned long long cand_cap;
  unsigned long long* n_cand;
  unsigned long long* n_bloom_pass;
  int W;
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
// (32 < W <= 64, two key words).  WN: number of competing m-mers.
template <int KW, int WN>
__global__ void __launch_bounds__(kScanBlock, MSC_SCAN_CTAS) scan_targets_kernel(const ScanArgs a) {
  pdl_enter();
  __shared__ alignas(128) uint64_t tiles[kScanWarps][2][kWarpSmemWords];
  __shared__ alignas(8) uint64_t bars[kScanWarps][2];
  __shared__ uint16_t queue[kScanWarps][1024];
  __shared__ uint2 stage[kScanWarps][kStageCap];  // found (slot, position) pairs, flushed when nearly full
  __shared__ uint32_t pattern[1024];              // bloom_pattern(): the two low-half bits of a key
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

  constexpr bool K32 = KW == 0;
  constexpr bool WIDE = KW == 2;
  const uint64_t kmask = low_bases_mask(min(a.W, 32));
  const uint64_t kmask1 = WIDE ? low_bases_mask(a.W - 32) : 0ull;  // bases 32..W-1 of a wide window
  const uint32_t xr = a.geom.xr;
  const int gm = a.geom.m;
  const int lg_words = a.geom.lg_words;
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
  auto flush_stage = [&]() {
    unsigned long long out0 = 0;
    if (lane == 0) out0 = atomicAdd(a.n_cand, (unsigned long long)n_st);
    out0 = __shfl_sync(0xffffffffu, out0, 0);
    // The slot's record travels with the candidate: it sits in the bucket line the look-up has just
    // brought into the L2, so this is the ONE dependent access a candidate costs (the expansion
    // reads no table memory at all).
#pragma unroll 4
    for (uint32_t i = lane; i < n_st; i += 32) {
      const uint2 e = st[i];
      const uint4 rec = __ldg(slot_rec_ptr(a.tab, (uint64_t)e.x));
      if (out0 + i < a.cand_cap) {
        a.cand[out0 + i] = e;
        a.cmeta[out0 + i] = rec;
      }
    }
    __syncwarp();
    n_st = 0;
  };
  for (; w0 < w_end; w0 += kWarpTileWords) {
    const uint64_t wn = w0 + kWarpTileWords;
    if (lane == 0 && wn < w_end) {
      mbar_arrive_expect_tx(&bar[buf ^ 1], kBytes);
      bulk_copy_g2s(tiles[warp][buf ^ 1], a.tg_words + wn, kBytes, &bar[buf ^ 1]);
    }
    const int tile_words = (int)min((uint64_t)kWarpTileWords, w_end - w0);  // multiple of kProbeBatch
    mbar_wait(&bar[buf], (phases >> buf) & 1u);
    phases ^= 1u << buf;
    const uint64_t* tile = tiles[warp][buf];

    // lane i owns word i of the tile: its pass mask and its X flag
    uint32_t mask = 0;
    // X summary bits of word w and w+1 (xsum is padded): bit i of xwords = word i of the tile needs the X path
    unsigned xwords;
    {
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
          const uint32_t sec = bloom_sector_of(bloom_min_mmer<WN>(prex, a.mul), gm, lg_words);
          bw[i] = __ldg(a.bloom + __funnelshift_l(h[i], sec, 2));  // (sec << 2) | (h >> 30)
        }
#pragma unroll
        for (int i = 0; i < kProbeBatch; i++) {
          // same masks as bloom_masks32(): pattern table look-up + rotate
          const uint32_t mlo = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(pattern) + (h[i] & 0xFFCu));
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
      // kDrain table look-ups in flight per lane: the home bucket (four fingerprints, one 256-bit
      // load) of each is fetched before any is resolved -- the look-ups are independent and a
      // single one costs an L2 / HBM round trip.
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
            const unsigned src = e[u] >> 5, j = e[u] & 31u;
            uint64_t xm = 0, xm1 = 0;
            const bool hasx = (xwords_all >> src) & 1u;
            if (hasx) xm = window_at(__ldg(a.tg_x + w0 + src), __ldg(a.tg_x + w0 + src + 1), j, kmask);
            if (WIDE) {
              if (hasx) xm1 = window_at(__ldg(a.tg_x + w0 + src + 1), __ldg(a.tg_x + w0 + src + 2), j, kmask1);
              fp[u] = key_fp_wide(window_at(tile[src], tile[src + 1], j, kmask), window_at(tile[src + 1], tile[src + 2], j, kmask1),
                                  xm, xm1);
            } else {
              fp[u] = key_fp(window_at(tile[src], tile[src + 1], j, kmask), xm);
            }
            bk[u] = table_home_bucket(fp[u], a.n_buckets);
            ldg256(bucket_ptr(a.tab, bk[u]), q[u][0], q[u][1], q[u][2], q[u][3]);  // the first four fingerprints of the home bucket
          }
        }
#pragma unroll
        for (int u = 0; u < kDrain; u++) {
          if (base + 32 * u >= total) break;  // warp-uniform
          // slots fill in order, so the fifth fingerprint only matters when the first four are taken
          // by other keys (rare at the table's load factor): it is fetched on demand
          int r = 5;  // 0..4 found, 5 = not in the table
          if (fp[u]) {
            while (true) {
              if (q[u][0] == fp[u]) { r = 0; break; }
              if (q[u][1] == fp[u]) { r = 1; break; }
              if (q[u][2] == fp[u]) { r = 2; break; }
              if (q[u][3] == fp[u]) { r = 3; break; }
              if ((q[u][0] == 0ull) | (q[u][1] == 0ull) | (q[u][2] == 0ull) | (q[u][3] == 0ull)) break;
              const uint64_t q4 = __ldg(reinterpret_cast<const unsigned long long*>(bucket_ptr(a.tab, bk[u]) + 32));
              if (q4 == fp[u]) { r = 4; break; }
              if (q4 == 0ull) break;
              bk[u] = bk[u] + 1 == a.n_buckets ? 0 : bk[u] + 1;  // bucket full of other keys: walk on
              ldg256(bucket_ptr(a.tab, bk[u]), q[u][0], q[u][1], q[u][2], q[u][3]);
            }
          }
          const bool hit = r < kBucketSlots;
          const unsigned found = __ballot_sync(0xffffffffu, hit);
          if (found) {
            if (hit)
              st[n_st + __popc(found & ((1u << lane) - 1u))] =
                  make_uint2((uint32_t)(bk[u] * kBucketSlots + (uint64_t)r), (uint32_t)(wbase + e[u]));
            n_st += __popc(found);
            __syncwarp();
            if (n_st > kStageCap - 32) flush_stage();
          }
        }
      }
      __syncwarp();
    }
    __syncwarp();  // all lanes are done with tile[buf] before lane 0 lets the TMA engine refill it
    buf ^= 1;
  }
  if (n_st) flush_stage();
  n_pass = __reduce_add_sync(0xffffffffu, n_pass);
  if (lane == 0 && n_pass) atomicAdd(a.n_bloom_pass, (unsigned long long)n_pass);
}

}  // namespace msc
