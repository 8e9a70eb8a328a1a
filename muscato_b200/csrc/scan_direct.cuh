// scan_direct.cuh -- kernel (2), exact-front variant: the target scan for W <= 15 against a key set so dense that a
// Bloom front stops filtering (BASELINE configs[2]: 3e8 keys of the 4^15 possible 15-mers, 28 % of all target
// positions are true hits).
//
// Replaces the same reference code as scan.cuh (processSeq / checkWin / harvest, cmd/muscato_screen/main.go:220-403).
// The front is ONE BIT PER POSSIBLE KEY (4^W bits, 128 MB at W = 15; BloomGeom::direct in common.cuh): bit x = key x
// is in the table.  No hash, no false positives for X-free windows -- but 128 MB does not stay in a 126 MB L2 that
// 77 GB of table lines stream through, and a probe that misses costs a whole HBM line.  So the scan runs in
// 2^lg_pass LAUNCHES; launch q walks the whole packed database (0.5 GB: nothing) and tests only the positions whose
// key falls into slice q of the bitmap (the key's top lg_pass bits = the window's last bases), so that the slice
// (16 MB) is L2-resident (createpolicy evict_last on the probes, evict_first on the table lines' last use and on the
// candidate records).  Measured at configs[2] on one B200: Bloom front 59.3 ms -> exact front 42.7 ms with the first
// version of this kernel (profiles/r02/call18_summary.txt).
//
// Per warp and 32-word tile (1024 positions, staged by the TMA engine as in scan.cuh):
//   members   lane i finds the positions of word i that belong to slice q with a few 64-bit operations on the word
//             pair (bit 2j of M <-> position j) and queues them -- a launch that tested every position would leave 1
//             lane in 2^lg_pass with work;
//   front     the queue is tested 32 x kFrontBatch entries at a time, every lane busy; survivors go to a SECOND queue
//             as (position, key, target) that outlives the tile;
//   drain     whenever that queue holds 64 entries: two table look-ups per lane, the candidate slots reserved with
//             one atomic that is in flight together with the bucket lines, records written straight to global memory
//             (an X-free survivor is in the table for certain, so nothing needs compacting; the rare false positive
//             -- an X window's hashed bit -- becomes a size-0 candidate, which the expansion skips).
// A tile does not wait for its own look-ups: with 16 slices a tile yields ~18 hits, and draining those alone left the
// warp on one HBM round trip per tile and launch (the first version's limit).
#pragma once
#include "scan.cuh"

namespace msc {

constexpr int kQB = 192;         // survivor queue entries per warp: < 64 carried over + <= 128 from one front round
constexpr int kFrontBatch = 4;   // front probes in flight per lane
constexpr int kDrainEntries = 64;

struct ScanDirectSmem {
  alignas(128) uint64_t tiles[kScanWarps][2][kWarpSmemWords];
  alignas(8) uint64_t bars[kScanWarps][2];
  uint16_t queue[kScanWarps][1024];  // members of the current tile: word << 5 | base
  uint32_t qb_pos[kScanWarps][kQB];  // survivors of the front: global position,
  uint32_t qb_key[kScanWarps][kQB];  //   key,
  uint32_t qb_g[kScanWarps][kQB];    //   target index
  uint32_t gtab[kScanWarps][kGeneTab + 4];
};

__global__ void __launch_bounds__(kScanBlock, MSC_SCAN_CTAS) scan_direct_kernel(const ScanArgs a) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char scan_smem[];
  ScanDirectSmem& sm = *reinterpret_cast<ScanDirectSmem*>(scan_smem);
  const int tid = threadIdx.x;
  const unsigned lane = tid & 31u, warp = tid >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  uint32_t* gt = sm.gtab[warp];
  uint16_t* q16 = sm.queue[warp];
  uint32_t* qb_pos = sm.qb_pos[warp];
  uint32_t* qb_key = sm.qb_key[warp];
  uint32_t* qb_g = sm.qb_g[warp];
  constexpr uint32_t kBytes = kWarpCopyWords * sizeof(uint64_t);
  uint64_t* bar = sm.bars[warp];
  if (lane == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  __syncwarp();

  const bool any_x = *a.targets_have_x != 0ull;
  const uint64_t kmask = low_bases_mask(a.W);  // W <= 15
  const uint32_t* bits32 = reinterpret_cast<const uint32_t*>(a.bloom);
  const unsigned pshift = (unsigned)(a.geom.lg_words + 6 - a.geom.lg_pass);  // slice of bit index x: x >> pshift
  const uint32_t my_pass = (uint32_t)a.pass;
  const uint64_t pol_keep = l2_policy_evict_last(), pol_once = l2_policy_evict_first();
  const bool stream_tab = a.stream_tab != 0;

  // the same static split of the stream as scan_targets_kernel
  const uint64_t n_words = (a.n_bases + 31) >> 5;
  const uint64_t n_units = (n_words + kProbeBatch - 1) / kProbeBatch;
  const uint64_t n_warps = (uint64_t)kScanWarps * gridDim.x, gw = (uint64_t)blockIdx.x * kScanWarps + warp;
  uint64_t w0 = (gw * n_units / n_warps) * kProbeBatch;
  const uint64_t w_end = ((gw + 1) * n_units / n_warps) * kProbeBatch;
  if (lane == 0 && w0 < w_end) {
    mbar_arrive_expect_tx(&bar[0], kBytes);
    bulk_copy_g2s(sm.tiles[warp][0], a.tg_words + w0, kBytes, &bar[0]);
  }
  uint32_t phases = 0;
  int buf = 0;
  uint32_t n_pass = 0;  // survivors of the front (lane 0 counts)
  uint32_t nB = 0;      // entries in the survivor queue (warp-uniform)
  uint32_t g0 = 0;      // target that holds the first base of the tile's first 2^kGeneBlockShift block
  bool gt_ok = false;   // gt[] covers the whole tile
  uint32_t gb1 = 0, gb2 = 0, gb3 = 0;

  // Target that holds position `pos` of the CURRENT tile (see scan.cuh).
  auto target_of = [&](uint32_t pos) -> uint32_t {
    if (pos < gb3) return g0 + (pos >= gb1) + (pos >= gb2);
    if (gt_ok) {
      uint32_t c = 0;
#pragma unroll
      for (int step = kGeneTab / 2; step >= 1; step >>= 1) c += (gt[c + step] <= pos) ? step : 0;
      return g0 + c;
    }
    const uint64_t g_lo = __ldg(a.blk2gene + (pos >> kGeneBlockShift)), g_hi = __ldg(a.blk2gene + (pos >> kGeneBlockShift) + 1);
    return (uint32_t)(upper_bound_dev<uint32_t>(a.tg_off, g_lo + 1, g_hi + 1, pos) - 1);
  };

  // One candidate, as the pair kernel wants it (scan.cuh, flush_stage): (slot, position), the key group's size, and
  // the 32-byte record.  hit == false: a reserved slot that found no key (size 0).
  auto emit = [&](unsigned long long o, bool hit, uint32_t slot, uint32_t pos, uint32_t g, uint32_t goff, uint32_t gend, uint4 rec) {
    if (o >= a.cand_cap) return;
    const uint32_t sz = (hit && (uint64_t)pos + (uint64_t)a.W <= (uint64_t)gend) ? 1u + rec.w : 0u;
    const uint2 e = make_uint2(slot, pos);
    const uint4 c0 = make_uint4(pos, pos - goff, gend, rec.y), c1 = make_uint4(rec.x, rec.z, g, 0u);
    if (stream_tab) {
      stg64_hint(a.cand + o, e, pol_once);
      stg128_hint(a.cinfo + 2 * o, c0, pol_once);
      stg128_hint(a.cinfo + 2 * o + 1, c1, pol_once);
      stg32_hint(a.sizes + o, sz, pol_once);
    } else {
      a.cand[o] = e;
      a.cinfo[2 * o] = c0;
      a.cinfo[2 * o + 1] = c1;
      a.sizes[o] = sz;
    }
  };

  // Drain the last cnt <= 64 entries of the survivor queue.
  auto drain = [&](uint32_t cnt) {
    const uint32_t first = nB - cnt;
    unsigned long long out0 = 0;
    if (lane == 0) out0 = atomicAdd(a.n_cand, (unsigned long long)cnt);  // in flight together with the bucket lines
    constexpr int kD = kDrainEntries / 32;
    uint64_t fp[kD], bk[kD], q[kD][4];
    uint32_t pos[kD], g[kD], goff[kD], gend[kD];
    bool valid[kD];
#pragma unroll
    for (int u = 0; u < kD; u++) {
      const uint32_t idx = 32u * u + lane;
      valid[u] = idx < cnt;
      fp[u] = 0;
      bk[u] = 0;
      pos[u] = g[u] = goff[u] = gend[u] = 0;
      q[u][0] = q[u][1] = q[u][2] = q[u][3] = 0;
      if (valid[u]) {
        fp[u] = (uint64_t)qb_key[first + idx] + 1ull;  // key_fp of an X-free window
        pos[u] = qb_pos[first + idx];
        g[u] = qb_g[first + idx];
        bk[u] = table_home_bucket(fp[u], a.n_buckets);
        ldg256(bucket_ptr(a.tab, bk[u]), q[u][0], q[u][1], q[u][2], q[u][3]);
        goff[u] = __ldg(a.tg_off + g[u]);
        gend[u] = __ldg(a.tg_off + g[u] + 1);
      }
    }
    int r[kD];
#pragma unroll
    for (int u = 0; u < kD; u++) {
      r[u] = 5;  // 0..4 found, 5 = not in the table
      if (valid[u]) {
        while (true) {
          if (q[u][0] == fp[u]) { r[u] = 0; break; }
          if (q[u][1] == fp[u]) { r[u] = 1; break; }
          if (q[u][2] == fp[u]) { r[u] = 2; break; }
          if (q[u][3] == fp[u]) { r[u] = 3; break; }
          if ((q[u][0] == 0ull) | (q[u][1] == 0ull) | (q[u][2] == 0ull) | (q[u][3] == 0ull)) break;
          const uint64_t q4 = __ldg(reinterpret_cast<const unsigned long long*>(bucket_ptr(a.tab, bk[u]) + 32));
          if (q4 == fp[u]) { r[u] = 4; break; }
          if (q4 == 0ull) break;
          bk[u] = bk[u] + 1 == a.n_buckets ? 0 : bk[u] + 1;
          ldg256(bucket_ptr(a.tab, bk[u]), q[u][0], q[u][1], q[u][2], q[u][3]);
        }
      }
    }
    uint4 rr[kD];
#pragma unroll
    for (int u = 0; u < kD; u++) {
      rr[u] = make_uint4(0u, 0u, 0u, 0u);
      if (r[u] < kBucketSlots) {
        const uint4* rp = reinterpret_cast<const uint4*>(bucket_ptr(a.tab, bk[u]) + kBucketRecOff) + r[u];
        rr[u] = stream_tab ? ldg128_last_use(rp, pol_once) : __ldg(rp);  // the line's last use: first out of the L2
      }
    }
    out0 = __shfl_sync(0xffffffffu, out0, 0);
#pragma unroll
    for (int u = 0; u < kD; u++)
      if (valid[u])
        emit(out0 + 32u * u + lane, r[u] < kBucketSlots, (uint32_t)(bk[u] * kBucketSlots + (uint64_t)(r[u] < kBucketSlots ? r[u] : 0)),
             pos[u], g[u], goff[u], gend[u], rr[u]);
    nB = first;
    __syncwarp();  // the queue tail has been read by every lane before the next front round writes there
  };

  uint32_t g0_next = w0 < w_end ? __ldg(a.blk2gene + ((w0 * 32ull) >> kGeneBlockShift)) : 0u;
  for (; w0 < w_end; w0 += kWarpTileWords) {
    const uint64_t wn = w0 + kWarpTileWords;
    if (lane == 0 && wn < w_end) {
      mbar_arrive_expect_tx(&bar[buf ^ 1], kBytes);
      bulk_copy_g2s(sm.tiles[warp][buf ^ 1], a.tg_words + wn, kBytes, &bar[buf ^ 1]);
    }
    const int tile_words = (int)min((uint64_t)kWarpTileWords, w_end - w0);
    g0 = g0_next;
    if (wn < w_end) g0_next = __ldg(a.blk2gene + ((wn * 32ull) >> kGeneBlockShift));
    const uint32_t gt_o0 = __ldg(a.tg_off + min((uint64_t)g0 + lane, a.n_targets));
    const uint32_t gt_o1 = __ldg(a.tg_off + min((uint64_t)g0 + 32u + lane, a.n_targets));
    const uint32_t gt_o2 = __ldg(a.tg_off + min((uint64_t)g0 + 64u, a.n_targets));
    mbar_wait(&bar[buf], (phases >> buf) & 1u);
    phases ^= 1u << buf;
    const uint64_t* tile = sm.tiles[warp][buf];
    const uint64_t wbase = w0 * 32ull;

    // words that contain X (or whose successor does): tested one by one with the X mask folded in (lane = position)
    unsigned xwords = 0;
    if (any_x) {
      const uint64_t w = w0 + (uint64_t)lane;
      const uint32_t xs0 = __ldg(a.xsum + (w >> 5));
      const uint32_t xs1 = __ldg(a.xsum + ((w + 1) >> 5));
      const uint32_t anyx = (xs0 >> (unsigned)(w & 31u)) | (xs1 >> (unsigned)((w + 1) & 31u));
      xwords = __ballot_sync(0xffffffffu, anyx & 1u);
    }
    if (tile_words < 32) xwords &= (1u << tile_words) - 1u;
    const unsigned xwords_all = xwords;
    uint32_t maskx = 0;  // lane i: survivors among the positions of X word i
    while (xwords) {
      const int wi = __ffs(xwords) - 1;
      xwords &= xwords - 1;
      const uint64_t xl = __ldg(a.tg_x + w0 + wi), xh = __ldg(a.tg_x + w0 + wi + 1);
      const uint64_t key = window_at(tile[wi], tile[wi + 1], lane, kmask), xm = window_at(xl, xh, lane, kmask);
      uint64_t widx;
      uint32_t mlo, mhi;
      bloom_locate(key, xm, xm ? key_fp(key, xm) : 0ull, a.W, a.geom, widx, mlo, mhi);
      const bool mine = (uint32_t)(widx >> (a.geom.lg_words - a.geom.lg_pass)) == my_pass;
      uint2 bwx = make_uint2(0u, 0u);
      if (mine) bwx = __ldg(a.bloom + widx);
      const unsigned b = __ballot_sync(0xffffffffu, mine & ((bwx.x & mlo) == mlo) & ((bwx.y & mhi) == mhi));
      if ((int)lane == wi) maskx = b;
    }

    // members of this launch's slice among the X-free words: bits [2p + pshift, 2p + 2W) of the stream spell my_pass
    uint64_t M = 0;
    {
      const uint64_t gbase = (w0 + (uint64_t)lane) * 32ull;
      if ((int)lane < tile_words && gbase < a.n_bases) {
        if (!((xwords_all >> lane) & 1u)) {
          const uint64_t lo = tile[lane], hi = tile[lane + 1];
          M = kEvenBits;
          for (int t = 0; t < a.geom.lg_pass; t++) {
            const unsigned sft = pshift + (unsigned)t;  // 6 <= sft < 2W <= 30
            const uint64_t S = (lo >> sft) | (hi << (64u - sft));
            M &= ((my_pass >> t) & 1u) ? S : ~S;
          }
        }
        if (a.n_bases - gbase < 32) {
          const unsigned nb = (unsigned)(a.n_bases - gbase);
          M &= (1ull << (2u * nb)) - 1ull;
          maskx &= (1u << nb) - 1u;
        }
      } else {
        maskx = 0;
      }
    }
    const uint32_t cntA = __popcll(M);
    uint32_t inclA = cntA;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, inclA, o);
      if ((int)lane >= o) inclA += v;
    }
    const uint32_t totalA = __shfl_sync(0xffffffffu, inclA, 31);
    const unsigned xsurv = __ballot_sync(0xffffffffu, maskx != 0u);
    if (totalA | xsurv) {
      // the tile's slice of the target offsets (scan.cuh)
      gt[lane] = gt_o0;
      gt[32 + lane] = gt_o1;
      if (lane == 0) gt[kGeneTab] = gt_o2;
      gt_ok = (uint64_t)gt_o2 > w0 * 32ull + (uint64_t)(32 * kWarpTileWords - 1);
      gb1 = __shfl_sync(0xffffffffu, gt_o0, 1);
      gb2 = __shfl_sync(0xffffffffu, gt_o0, 2);
      gb3 = __shfl_sync(0xffffffffu, gt_o0, 3);
      {
        uint32_t at = inclA - cntA;
        while (M) {
          const uint32_t j2 = __ffsll((long long)M) - 1;
          M &= M - 1;
          q16[at++] = (uint16_t)((lane << 5) | (j2 >> 1));
        }
      }
      __syncwarp();
      // front: 32 x kFrontBatch queued members per round, survivors to the cross-tile queue
      for (uint32_t base = 0; base < totalA; base += 32 * kFrontBatch) {
        uint32_t e[kFrontBatch], key[kFrontBatch], bwd[kFrontBatch];
#pragma unroll
        for (int u = 0; u < kFrontBatch; u++) {
          const uint32_t idx = base + 32 * u + lane;
          e[u] = idx < totalA ? q16[idx] : 0xffffffffu;
          key[u] = 0;
          bwd[u] = 0;
          if (e[u] != 0xffffffffu) {
            const unsigned src = e[u] >> 5, j = e[u] & 31u;
            key[u] = (uint32_t)window_at(tile[src], tile[src + 1], j, kmask);
            bwd[u] = ldg32_hint(bits32 + (key[u] >> 5), pol_keep);
          }
        }
#pragma unroll
        for (int u = 0; u < kFrontBatch; u++) {
          const bool hit = (bwd[u] >> (key[u] & 31u)) & 1u;
          const unsigned found = __ballot_sync(0xffffffffu, hit);
          if (hit) {
            const uint32_t at = nB + __popc(found & lt_mask), pos = (uint32_t)(wbase + e[u]);
            qb_pos[at] = pos;
            qb_key[at] = key[u];
            qb_g[at] = target_of(pos);
          }
          nB += __popc(found);
          if (lane == 0) n_pass += __popc(found);
        }
        __syncwarp();
        while (nB >= (uint32_t)kDrainEntries) drain(kDrainEntries);
      }
      // survivors of the X words (rare): looked up on the spot, lane = position
      unsigned xs = xsurv;
      while (xs) {
        const int wi = __ffs(xs) - 1;
        xs &= xs - 1;
        const uint32_t mw = __shfl_sync(0xffffffffu, maskx, wi);
        const bool mine = (mw >> lane) & 1u;
        int64_t slot = -1;
        uint4 rec = make_uint4(0u, 0u, 0u, 0u);
        if (mine) {
          const uint64_t key = window_at(tile[wi], tile[wi + 1], lane, kmask);
          const uint64_t xm = window_at(__ldg(a.tg_x + w0 + wi), __ldg(a.tg_x + w0 + wi + 1), lane, kmask);
          slot = table_find(a.tab, a.n_buckets, key_fp(key, xm));
          if (slot >= 0) rec = __ldg(slot_rec_ptr(a.tab, (uint64_t)slot));
        }
        const unsigned found = __ballot_sync(0xffffffffu, slot >= 0);
        if (lane == 0) n_pass += __popc(mw);
        if (found) {
          unsigned long long out0 = 0;
          if (lane == 0) out0 = atomicAdd(a.n_cand, (unsigned long long)__popc(found));
          out0 = __shfl_sync(0xffffffffu, out0, 0);
          if (slot >= 0) {
            const uint32_t pos = (uint32_t)(wbase + 32u * (unsigned)wi + lane), g = target_of(pos);
            emit(out0 + __popc(found & lt_mask), true, (uint32_t)slot, pos, g, __ldg(a.tg_off + g), __ldg(a.tg_off + g + 1), rec);
          }
        }
      }
    }
    __syncwarp();  // all lanes are done with tile[buf] before lane 0 lets the TMA engine refill it
    buf ^= 1;
  }
  if (nB) drain(nB);
  if (lane == 0 && n_pass) atomicAdd(a.n_bloom_pass, (unsigned long long)n_pass);
}

}  // namespace msc
