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
// (32 MB) is L2-resident (createpolicy evict_last on the probes, evict_first on the table lines' last use and on the
// candidate records).  Measured at configs[2] on one B200: Bloom front 59.3 ms -> exact front 36.5 ms
// (profiles/r02/call16..19, call24, call34 summaries; what bounds it now: call26..29, call38).
//
// Per warp and 32-word tile (1024 positions, staged by the TMA engine as in scan.cuh):
//   members   lane i finds the positions of word i that belong to slice q with a few 64-bit operations on the word
//             pair (a launch that tested every position would leave 1 lane in 2^lg_pass with work) and keeps them as
//             a 32-bit mask, the word pair as three registers;
//   front     every lane walks ITS OWN members, kFrontBatch at a time: key = one funnel shift of two registers, one
//             4-byte probe of the slice (no shared-memory traffic at all: the first version queued the members and
//             re-read the tile for every one of them -- 45 % of its instructions and its main stall);
//             survivors go to a queue of (position, key, target) that outlives the tile;
//   drain     whenever that queue holds 64 entries: two table look-ups per lane; candidate slots come out of a block
//             of kSlotBlock that the warp reserves with ONE atomic (requested while the bucket lines are in flight),
//             records are written straight to global memory (an X-free survivor is in the table for certain, so
//             nothing needs compacting; the rare false positive -- an X window's hashed bit -- and the unused tail
//             of a warp's last block become size-0 candidates, which the expansion skips).
// A tile does not wait for its own look-ups: with 16 slices a tile yields ~18 hits, and draining those alone left the
// warp on one HBM round trip per tile and launch.
#pragma once
#include "scan.cuh"

namespace msc {

#ifndef MSC_DRAIN_ENTRIES
#define MSC_DRAIN_ENTRIES 64
#endif
#ifndef MSC_DIRECT_CTAS
#define MSC_DIRECT_CTAS 3  // 74 registers, no spills; measured 36.7 ms at configs[2] against 37.8 ms with four CTAs of 64 (spilling) registers
#endif
constexpr int kFrontBatch = 4;   // front probes in flight per lane
constexpr int kDrainEntries = MSC_DRAIN_ENTRIES;  // look-ups per drain (32 per lane-round)
constexpr int kQB = kDrainEntries + 32 * kFrontBatch;  // survivor queue entries per warp: < one drain carried over + one front round
constexpr int kSlotBlock = 256;  // candidate slots a warp reserves per atomic
constexpr int kDBufs = 4;        // tile buffers per warp: three TMA copies in flight (a tile-pass with few members is
                                 // shorter than one HBM round trip)
constexpr int kDGeneTab = 32;    // target offsets a warp keeps in shared memory per tile

struct ScanDirectSmem {
  alignas(128) uint64_t tiles[kScanWarps][kDBufs][kWarpSmemWords];
  alignas(8) uint64_t bars[kScanWarps][kDBufs];
  uint32_t qb_pos[kScanWarps][kQB];  // survivors of the front: global position,
  uint32_t qb_key[kScanWarps][kQB];  //   key,
  uint32_t qb_g[kScanWarps][kQB];    //   target index,
  uint32_t qb_bk[kScanWarps][kQB];   //   home bucket in the key table
  uint32_t gtab[kScanWarps][kDGeneTab + 4];
};

__global__ void __launch_bounds__(kScanBlock, MSC_DIRECT_CTAS) scan_direct_kernel(const ScanArgs a) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char scan_smem[];
  ScanDirectSmem& sm = *reinterpret_cast<ScanDirectSmem*>(scan_smem);
  const int tid = threadIdx.x;
  const unsigned lane = tid & 31u, warp = tid >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  uint32_t* gt = sm.gtab[warp];
  uint32_t* qb_pos = sm.qb_pos[warp];
  uint32_t* qb_key = sm.qb_key[warp];
  uint32_t* qb_g = sm.qb_g[warp];
  uint32_t* qb_bk = sm.qb_bk[warp];
  constexpr uint32_t kBytes = kWarpCopyWords * sizeof(uint64_t);
  uint64_t* bar = sm.bars[warp];
  if (lane == 0) {
    for (int b = 0; b < kDBufs; b++) mbar_init(&bar[b], 1);
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
  if (lane == 0) {
    for (int b = 0; b < kDBufs - 1; b++) {
      const uint64_t wt = w0 + (uint64_t)b * kWarpTileWords;
      if (wt < w_end) {
        mbar_arrive_expect_tx(&bar[b], kBytes);
        bulk_copy_g2s(sm.tiles[warp][b], a.tg_words + wt, kBytes, &bar[b]);
      }
    }
  }
  uint32_t phases = 0;
  int buf = 0;
  uint32_t n_pass = 0;  // survivors of the front (lane 0 counts)
  uint32_t n_dummy = 0; // reserved candidate slots that stayed empty (lane 0 counts)
  unsigned long long slot_base = 0;  // the warp's block of reserved candidate slots: next free slot,
  uint32_t slot_left = 0;            //   slots left (both warp-uniform)
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
      for (int step = kDGeneTab / 2; step >= 1; step >>= 1) c += (gt[c + step] <= pos) ? step : 0;
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
    const uint4 c0 = make_uint4(pos, pos - goff, gend, rec.y), c1 = make_uint4(rec.x, rec.z, g, slot);
    if (stream_tab) {
      stg128_hint(a.cinfo + 2 * o, c0, pol_once);
      stg128_hint(a.cinfo + 2 * o + 1, c1, pol_once);
      stg32_hint(a.sizes + o, sz, pol_once);
    } else {
      a.cinfo[2 * o] = c0;
      a.cinfo[2 * o + 1] = c1;
      a.sizes[o] = sz;
    }
  };

  // Empty candidates for reserved slots [o0, o0 + n) that nothing will use (n < kSlotBlock).
  auto pad_slots = [&](unsigned long long o0, uint32_t n) {
    for (uint32_t i = lane; i < n; i += 32)
      if (o0 + i < a.cand_cap) a.sizes[o0 + i] = 0u;
    if (lane == 0) n_dummy += n;
  };
  // cnt candidate slots for the warp (cnt <= kDrainEntries): from its block, or from a new block when the block is too
  // small (its tail is padded).  The atomic's result is only touched by take_slots_finish().
  unsigned long long new_block = 0;
  bool block_pending = false;
  auto take_slots_begin = [&](uint32_t cnt) {
    block_pending = slot_left < cnt;
    if (block_pending) {
      if (slot_left) pad_slots(slot_base, slot_left);
      if (lane == 0) new_block = atomicAdd(a.n_cand, (unsigned long long)kSlotBlock);
    }
  };
  auto take_slots_finish = [&](uint32_t cnt) -> unsigned long long {
    if (block_pending) {
      slot_base = __shfl_sync(0xffffffffu, new_block, 0);
      slot_left = kSlotBlock;
    }
    const unsigned long long o = slot_base;
    slot_base += cnt;
    slot_left -= cnt;
    return o;
  };

  // Drain the last cnt <= 64 entries of the survivor queue.
  auto drain = [&](uint32_t cnt) {
    const uint32_t first = nB - cnt;
    take_slots_begin(cnt);  // (a new block's atomic is in flight together with the bucket lines)
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
        bk[u] = qb_bk[first + idx];
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
    const unsigned long long out0 = take_slots_finish(cnt);
#pragma unroll
    for (int u = 0; u < kD; u++)
      if (valid[u])
        emit(out0 + 32u * u + lane, r[u] < kBucketSlots, (uint32_t)(bk[u] * kBucketSlots + (uint64_t)(r[u] < kBucketSlots ? r[u] : 0)),
             pos[u], g[u], goff[u], gend[u], rr[u]);
#pragma unroll
    for (int u = 0; u < kD; u++) {
      const unsigned miss = __ballot_sync(0xffffffffu, valid[u] && r[u] >= kBucketSlots);
      if (lane == 0) n_dummy += __popc(miss);
    }
    nB = first;
    __syncwarp();  // the queue tail has been read by every lane before the next front round writes there
  };

  // The tile's slice of the target offsets is requested ONE TILE AHEAD (tile t asks for tile t + 1's offsets and for
  // tile t + 2's first target), so that nothing of it is waited for inside a tile.
  auto first_target = [&](uint64_t w) -> uint32_t { return __ldg(a.blk2gene + ((w * 32ull) >> kGeneBlockShift)); };
  uint32_t g0_n = 0, g0_nn = 0, gt_o0 = 0, gt_o1 = 0;  // next tile's first target, the one after; this tile's offsets
  if (w0 < w_end) {
    g0_n = first_target(w0);
    if (w0 + kWarpTileWords < w_end) g0_nn = first_target(w0 + kWarpTileWords);
    gt_o0 = __ldg(a.tg_off + min((uint64_t)g0_n + lane, a.n_targets));
    gt_o1 = __ldg(a.tg_off + min((uint64_t)g0_n + (uint64_t)kDGeneTab, a.n_targets));
  }
  for (; w0 < w_end; w0 += kWarpTileWords) {
    const uint64_t wn = w0 + kWarpTileWords;
    {
      const uint64_t wt = w0 + (uint64_t)(kDBufs - 1) * kWarpTileWords;  // the buffer the previous tile has just left
      const int bt = (buf + kDBufs - 1) % kDBufs;
      if (lane == 0 && wt < w_end) {
        mbar_arrive_expect_tx(&bar[bt], kBytes);
        bulk_copy_g2s(sm.tiles[warp][bt], a.tg_words + wt, kBytes, &bar[bt]);
      }
    }
    const int tile_words = (int)min((uint64_t)kWarpTileWords, w_end - w0);
    g0 = g0_n;
    g0_n = g0_nn;
    const uint32_t gt_c0 = gt_o0, gt_c1 = gt_o1;  // this tile's
    if (wn < w_end) {
      gt_o0 = __ldg(a.tg_off + min((uint64_t)g0_n + lane, a.n_targets));
      gt_o1 = __ldg(a.tg_off + min((uint64_t)g0_n + (uint64_t)kDGeneTab, a.n_targets));
      if (wn + kWarpTileWords < w_end) g0_nn = first_target(wn + kWarpTileWords);
    }
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

    // members of this launch's slice among the X-free words: bits [2p + pshift, 2p + 2W) of the stream spell my_pass.
    // Mp: bit j = position j of word `lane` is a member; (w0r, w1r, w2r) = the 96 stream bits its keys come from.
    uint32_t Mp = 0, w0r = 0, w1r = 0, w2r = 0;
    {
      const uint64_t gbase = (w0 + (uint64_t)lane) * 32ull;
      if ((int)lane < tile_words && gbase < a.n_bases) {
        if (!((xwords_all >> lane) & 1u)) {
          const uint64_t lo = tile[lane], hi = tile[lane + 1];
          uint64_t M = kEvenBits;
          for (int t = 0; t < a.geom.lg_pass; t++) {
            const unsigned sft = pshift + (unsigned)t;  // 6 <= sft < 2W <= 30
            const uint64_t S = (lo >> sft) | (hi << (64u - sft));
            M &= ((my_pass >> t) & 1u) ? S : ~S;
          }
          Mp = compress_even32((uint32_t)M) | (compress_even32((uint32_t)(M >> 32)) << 16);
          w0r = (uint32_t)lo;
          w1r = (uint32_t)(lo >> 32);
          w2r = (uint32_t)hi;
        }
        if (a.n_bases - gbase < 32) {
          const unsigned nb = (unsigned)(a.n_bases - gbase);
          Mp &= (1u << nb) - 1u;
          maskx &= (1u << nb) - 1u;
        }
      } else {
        maskx = 0;
      }
    }
    const unsigned xsurv = __ballot_sync(0xffffffffu, maskx != 0u);
    if (__any_sync(0xffffffffu, Mp != 0u) | (xsurv != 0u)) {
      // the tile's slice of the target offsets (scan.cuh)
      gt[lane] = gt_c0;
      if (lane == 0) gt[kDGeneTab] = gt_c1;
      gt_ok = (uint64_t)gt_c1 > w0 * 32ull + (uint64_t)(32 * kWarpTileWords - 1);
      gb1 = __shfl_sync(0xffffffffu, gt_c0, 1);
      gb2 = __shfl_sync(0xffffffffu, gt_c0, 2);
      gb3 = __shfl_sync(0xffffffffu, gt_c0, 3);
      __syncwarp();
      // front: every lane tests kFrontBatch of its own members per round; survivors to the cross-tile queue
      const uint32_t wpos = (uint32_t)wbase + 32u * lane;  // first position of this lane's word
      while (__any_sync(0xffffffffu, Mp != 0u)) {
        uint32_t key[kFrontBatch], bwd[kFrontBatch], jj[kFrontBatch];
#pragma unroll
        for (int u = 0; u < kFrontBatch; u++) {
          const bool has = Mp != 0u;
          const uint32_t j = has ? (uint32_t)__ffs(Mp) - 1u : 0u;
          Mp &= Mp - 1u;  // (0 stays 0)
          // the 30 stream bits that start at bit 2j of (w2r:w1r:w0r)
          const uint32_t lo32 = j & 16u ? w1r : w0r, hi32 = j & 16u ? w2r : w1r;
          key[u] = __funnelshift_r(lo32, hi32, (2u * j) & 31u) & (uint32_t)kmask;
          jj[u] = j;
          bwd[u] = 0u;
          if (has) bwd[u] = ldg32_hint(bits32 + (key[u] >> 5), pol_keep);
        }
#pragma unroll
        for (int u = 0; u < kFrontBatch; u++) {
          const bool hit = (bwd[u] >> (key[u] & 31u)) & 1u;
          const unsigned found = __ballot_sync(0xffffffffu, hit);
          if (hit) {
            const uint32_t at = nB + __popc(found & lt_mask), pos = wpos + jj[u];
            qb_pos[at] = pos;
            qb_key[at] = key[u];
            qb_g[at] = target_of(pos);
            // (MSC_SCAN_PREFETCH=1/2 requests the survivor's bucket line into the L2 right here, by prefetch.global.L2 or
            // through the TMA engine: both were measured SLOWER -- 55 / 53 ms against 37 ms at configs[2] -- and are off)
            const uint64_t bk = table_home_bucket((uint64_t)key[u] + 1ull, a.n_buckets);
            qb_bk[at] = (uint32_t)bk;
            if (a.prefetch == 1) prefetch_l2(bucket_ptr(a.tab, bk));
            else if (a.prefetch == 2) bulk_prefetch_l2(bucket_ptr(a.tab, bk), kBucketBytes);
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
    buf = (buf + 1) % kDBufs;
  }
  if (nB) drain(nB);
  if (slot_left) pad_slots(slot_base, slot_left);  // the unused tail of the warp's last block
  if (lane == 0 && n_pass) atomicAdd(a.n_bloom_pass, (unsigned long long)n_pass);
  if (lane == 0 && n_dummy) atomicAdd(a.n_dummy, (unsigned long long)n_dummy);
}

}  // namespace msc
