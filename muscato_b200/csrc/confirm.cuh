// confirm.cuh -- kernels (3) and (4): segmented expansion of candidates into
// (candidate, read) pairs and the XOR+popcount confirm.
//
// (3) replaces sortBloom (cmd/muscato/main.go:318-385) and the merge join of
//     cmd/muscato_confirm/main.go:375-416: every candidate already carries its key-group
//     (table slot), so grouping is a prefix sum over the group sizes; pair i is located by
//     a binary search in that prefix array (load-balanced: one thread per pair, however
//     skewed the groups are).
// (4) replaces searchpairs/cdiff (cmd/muscato_confirm/main.go:151-250): full-read mismatch
//     count on packed words, the fit rule (:200-203), the position-0 rule of the screen
//     (cmd/muscato_screen/main.go:294-316, the literal 100), nmiss = int((1-PMatch)*L) from a
//     host-computed float64 table (:198), and exact cross-window de-duplication (the
//     `sort -u` of cmd/muscato/main.go:453-463): a pair is emitted only through the lowest
//     window index that delivers it.
#pragma once
#include "build.cuh"
#include "common.cuh"

namespace msc {

// Pairs are addressed in blocks of kPairBlock consecutive pair indices (block_first[]).
constexpr int kPairBlockShift = 5;
constexpr int kPairBlock = 1 << kPairBlockShift;

// The candidates arrive from the scan kernel complete (scan.cuh, flush_stage): (slot, position), the key group's
// size, and one 32-byte record per candidate -- (global position of the window, window start p inside the gene,
// global end of the gene, read record word of the first item) and (first item of the key group, CSR start of the
// further items, gene index, 0).  The expansion is therefore the exclusive scan of the sizes plus the block index below.

// First candidate of every kPairBlock-pair block of the confirm kernel, so that the search inside
// the kernel only spans the block's few candidates.  Entry n_blocks is a sentinel.  block_first[b]
// is the candidate whose pair range [pstart[c], pstart[c+1]) contains the block's first pair
// (the last pair for the sentinel of a complete list): instead of one binary search per block, every
// candidate writes the entries of the blocks that start inside its range (coalesced reads of
// pstart, ~n_blocks scattered 4-byte stores in total).
__global__ void __launch_bounds__(256) pair_block_starts_kernel(const uint64_t* __restrict__ pstart,
                                                                const unsigned long long* __restrict__ n_cand_ptr,
                                                                uint64_t cand_cap,
                                                                const unsigned long long* __restrict__ n_pairs_ptr,
                                                                uint64_t block_cap, uint32_t* __restrict__ block_first) {
  pdl_enter();
  const uint64_t n_cand = min((uint64_t)*n_cand_ptr, cand_cap);
  const uint64_t n_pairs = *n_pairs_ptr;
  if (n_pairs == 0) return;
  const uint64_t n_blocks = min((uint64_t)((n_pairs + kPairBlock - 1) >> kPairBlockShift), block_cap);
  const bool complete = (n_blocks << kPairBlockShift) >= n_pairs;  // false while block_first is still too small
  for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_cand; c += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t s = __ldg(pstart + c), e = __ldg(pstart + c + 1);
    if (e <= s) continue;  // no pairs
    const uint64_t b1 = min((uint64_t)((e - 1) >> kPairBlockShift), n_blocks);
    for (uint64_t b = (s + kPairBlock - 1) >> kPairBlockShift; b <= b1; b++) block_first[b] = (uint32_t)c;
    if (complete && e == n_pairs) block_first[n_blocks] = (uint32_t)c;  // sentinel: the candidate of the last pair
  }
}

struct ConfirmArgs {
  // candidates and their pair prefix
  const uint4* cinfo;          // two per candidate, written by the scan kernel
  const uint32_t* block_first; // first candidate of each kPairBlock-pair block (+1 sentinel entry)
  const uint64_t* pstart;  // n_cand + 1
  const unsigned long long* n_pairs_ptr;  // device-side pair count (grand total of the size scan)
  uint64_t block_cap;                     // capacity of block_first (in kPairBlock-pair blocks)
  // key table
  const uint2* items;          // CSR of further group members: (item, read record word)
  const uint32_t* validmask;   // per read: valid-window mask (only read for pairs that pass through a window k > 0)
  // MaxMatches pre-check: passing pairs are counted per key group in a SMALL hashed counter array that
  // stays in the L2 (a counter is an upper bound for every group that maps to it: if none exceeds
  // MaxMatches no group does); the exact per-slot counts (pass_cnt != nullptr) are only taken in the
  // rare re-run that follows a counter above the limit
  uint32_t* pass_small;
  int lg_small;
  uint32_t* pass_cnt;
  // reads
  const uint64_t* rd_words;
  const uint64_t* rd_x;
  // targets
  const uint64_t* tg_words;
  const uint64_t* tg_x;
  const uint32_t* xsum;
  const uint32_t* tg_off;  // n_targets + 1 (base offsets in the concatenated stream)
  uint64_t n_targets;
  // rules
  uint64_t nwin_magic;   // floor(2^64 / nwin) + 1: item / nwin == umul64hi(item, magic) for 32-bit items
  const unsigned long long* targets_have_x;  // device flag: 0 = no target word contains X
  // outputs
  uint4* matches;  // (read, gene, pos, nx)
  unsigned long long match_cap;
  unsigned long long* n_match;
  unsigned long long* n_pass;
  uint32_t* best;  // per read min nx
  int mode;        // 0 = confirm, 1 = dump exact-key candidates as (gene, p, read, window),
                   // 2 = confirm with MaxMatches overflow groups diverted (see below)
  // mode 2 only: key groups whose passing-pair count exceeds MaxMatches are flagged in
  // slot_over.  A passing pair found through such a group is written to `over` as
  // (read, gene, pos, nx | window<<16) for the host-side emulation of the reference's
  // order-dependent truncation (cmd/muscato_confirm/main.go:233-242, :424-448); the
  // cross-window de-duplication only counts windows whose group is not flagged.
  const uint8_t* slot_over;
  const uint8_t* tab;
  uint64_t n_buckets;
  uint4* over;
  unsigned long long over_cap;
  unsigned long long* n_over_inst;
};

__device__ __forceinline__ bool tg_range_has_x(const uint32_t* __restrict__ xsum, uint64_t w0, uint64_t w1) {
  // any X summary bit set for words w0..w1 (inclusive)
  for (uint64_t u = w0 >> 5; u <= (w1 >> 5); u++) {
    uint32_t v = __ldg(xsum + u);
    if (!v) continue;
    const uint64_t lo = u << 5;
    if (w0 > lo) v &= ~0u << (unsigned)(w0 - lo);
    if (w1 < lo + 31) v &= ~0u >> (unsigned)(lo + 31 - w1);
    if (v) return true;
  }
  return false;
}

// 32 bases of a packed stream starting at bit offset sh (0..62, even) of the word pair (t0, t1).
__device__ __forceinline__ uint64_t funnel64(uint64_t t0, uint64_t t1, unsigned sh) {
  return (t0 >> sh) | ((t1 << 1) << (63u - sh));
}

// Exact equality of a read window with the target window that lies on it (bases and X masks),
// in chunks of 32 bases (one chunk for W <= 32): the byte comparison of the merge join
// (cmd/muscato_confirm/main.go:382-393).  rq / tq = first base of the window in the read row /
// in the target stream.
__device__ __forceinline__ bool window_match(const uint64_t* __restrict__ row, const uint64_t* __restrict__ xrow, bool rx,
                                             const uint64_t* __restrict__ tg_words, const uint64_t* __restrict__ tg_x,
                                             bool tx, uint64_t rq, uint64_t tq, int W) {
  for (int off = 0; off < W; off += 32) {
    const uint64_t mk = low_bases_mask(min(32, W - off));
    if ((extract32(row, rq + off) & mk) != (extract32(tg_words, tq + off) & mk)) return false;
    if (rx | tx) {
      const uint64_t rxm = rx ? (extract32(xrow, rq + off) & mk) : 0ull;
      const uint64_t txm = tx ? (extract32(tg_x, tq + off) & mk) : 0ull;
      if (rxm != txm) return false;
    }
  }
  return true;
}

// Fingerprint of window [rq, rq + W) of a read, as the table stores it (build_keys_insert_kernel).
__device__ __forceinline__ uint64_t read_window_fp(const uint64_t* __restrict__ row, const uint64_t* __restrict__ xrow,
                                                   bool rx, uint64_t rq, int W) {
  const uint64_t mk = low_bases_mask(min(W, 32));
  const uint64_t k0 = extract32(row, rq) & mk, x0 = rx ? (extract32(xrow, rq) & mk) : 0ull;
  if (W <= 32) return key_fp(k0, x0);
  const uint64_t mk1 = low_bases_mask(W - 32);
  return key_fp_wide(k0, extract32(row, rq + 32) & mk1, x0, rx ? (extract32(xrow, rq + 32) & mk1) : 0ull);
}

// One (candidate, read) pair; c = index of its candidate.  MODE is a compile-time copy of
// ConfirmArgs::mode so that the hot mode-0 kernel carries none of the tap / overflow code.
// Returns true when the pair yields an output record (`rec`): a match (read, gene, pos, nx) in
// modes 0/2, an exact-key candidate (gene, p, read, window) in mode 1.  n_pass counts pairs
// that passed through their window (before cross-window de-duplication).
//
// Fast path (no X in the read or in the target range, W < 32; wide windows, W > 32, always take the
// re-check): the table's fingerprints are
// exact there (common.cuh), so the window that produced the pair matches by construction and
// the pair only needs the fit rule and the full-read mismatch count -- target words are loaded
// once each (nwords + 1 loads) and funnel-shifted against the read's row.
template <int MODE>
__device__ __forceinline__ bool confirm_one_pair(const WinCfg& cfg, const ConfirmArgs& a, uint64_t i, uint64_t c,
                                                 uint64_t c_start, bool targets_have_x, uint4& rec,
                                                 uint32_t& n_pass) {
  // the candidate's 32-byte record in one 256-bit load:
  // (position, p, gene end, read record .x of item 0 | item 0, CSR start, gene, read record .y of item 0)
  uint4 ci, cj;
  {
    uint64_t q0, q1, q2, q3;
    ldg256(a.cinfo + 2 * c, q0, q1, q2, q3);
    ci = make_uint4((uint32_t)q0, (uint32_t)(q0 >> 32), (uint32_t)q1, (uint32_t)(q1 >> 32));
    cj = make_uint4((uint32_t)q2, (uint32_t)(q2 >> 32), (uint32_t)q3, (uint32_t)(q3 >> 32));
  }
  const uint64_t gpos = ci.x;
  const int64_t p = (int64_t)ci.y;
  const uint32_t gend = ci.z;
  const uint32_t j = (uint32_t)(i - c_start);
  // (item, read record): member 0 comes with the candidate record, further members from the CSR
  uint32_t item = cj.x;
  uint32_t rmx = ci.w;
  if (j) {
    const uint2 e = __ldg(a.items + cj.y + (j - 1));
    item = e.x;
    rmx = e.y;
  }
  const uint32_t r = cfg.nwin == 1 ? item : (uint32_t)__umul64hi((uint64_t)item, a.nwin_magic);  // item / nwin
  const int k = (int)(item - r * (uint32_t)cfg.nwin);
  const int W = cfg.W;
  const int q1 = cfg.windows[k];
  const int64_t pos = p - q1;      // jw = jx - q1 >= 0 (cmd/muscato_screen/main.go:345, :355)
  if (pos < 0) return false;

  const int L = rmx_len(cfg, rmx);
  const int budget = rmx_budget(cfg, rmx);
  const bool rx = rmx >> 31;
  const uint64_t* row = a.rd_words + (uint64_t)r * cfg.S;
  const uint64_t gstart = gpos - (uint64_t)q1;  // global base index of the read's first base
  const int64_t glen = (int64_t)gend - (int64_t)(gpos - (uint64_t)p);
  const int64_t lim0 = min((int64_t)(100 - W), glen);  // position-0 record: right = t[W : min(100-q2, len)]

  if (MODE != 1) {
    // Fit rule (cmd/muscato_confirm/main.go:200-203) on the candidate's clipped right tail:
    // for p >= 1 it reduces to pos + L <= len(target) (L <= MaxReadLength always holds); the
    // position-0 record carries right = t[W : min(100 - q2, len)] (the literal 100, Q1).
    if (p == 0) {
      if ((int64_t)L > lim0) return false;
    } else if (gstart + (uint64_t)L > (uint64_t)gend) {
      return false;
    }
  }

  const bool tx = targets_have_x && tg_range_has_x(a.xsum, gstart >> 5, ((gstart + (uint64_t)L) >> 5) + 1);
  const bool anyx = rx | tx;
  const uint64_t* xrow = a.rd_x + (uint64_t)r * cfg.S;

  // Sketch pre-filter: the item carries sk read bases next to its window (build.cuh, sketch_start); the target bases
  // they would lie on are in lines the neighbouring candidates share.  More mismatches there than the budget allows
  // rejects the pair (exact: the sketch bases are part of the full compare) WITHOUT touching the read's row -- at
  // scale a row costs one 128-byte HBM line per pair and most pairs of a short window are chance hits.
  if (MODE != 1 && !anyx) {
    const int s0 = sketch_start(cfg, k, L);
    if (s0 >= 0) {
      const uint32_t tb = (uint32_t)(extract32(a.tg_words, gstart + (uint64_t)s0) & low_bases_mask(cfg.sk));
      const uint32_t x = tb ^ rmx_sketch(cfg, rmx);
      if (__popc((x | (x >> 1)) & 0x55555555u) > budget) return false;
    }
  }

  // Exact key equality for the window that produced this pair (merge join on the k-mer bytes,
  // cmd/muscato_confirm/main.go:382-393).  Fingerprints of X-free W<32 windows are exact, so only
  // pairs that involve X (or W >= 32) can be fingerprint collisions.
  if (MODE == 1 || anyx || W >= 32) {
    if (!window_match(row, xrow, rx, a.tg_words, a.tg_x, tx, (uint64_t)q1, gpos, W)) return false;
    if (MODE == 1) {
      rec = make_uint4(cj.z, (uint32_t)p, r, (uint32_t)k);
      return true;
    }
  }

  // Full-read mismatch count: nx = cdiff(left tails) + cdiff(right tails) (+0 inside the window).
  // Early exit once the budget is exceeded (the count itself is only needed for kept pairs).
  int nx = 0;
  {
    const int nwords = (L + 31) >> 5;
    const unsigned sh = (unsigned)(gstart & 31u) * 2u;
    const uint64_t* tw = a.tg_words + (gstart >> 5);
    const uint64_t* txw = a.tg_x + (gstart >> 5);
    if (!anyx && nwords <= 4) {
      // Short reads without X (the norm): all loads are issued before the first compare, so the
      // early exit costs no dependent round trips.
      uint64_t rw[4], t[5];
      if ((cfg.S & 1) == 0) {
        // 16-byte loads: rows of an even number of words are 16-byte aligned, and the 5 target
        // words lie inside the 6 words that start at the even word index below them (buffers are
        // padded): 5 requests per pair instead of 9
        if ((cfg.S & 3) == 0) {  // 32-byte aligned rows: one 256-bit load
          ldg256(row, rw[0], rw[1], rw[2], rw[3]);
        } else {
          const ulonglong2* r2 = reinterpret_cast<const ulonglong2*>(row);
          const ulonglong2 ra = __ldg(r2), rb = nwords > 2 ? __ldg(r2 + 1) : make_ulonglong2(0ull, 0ull);
          rw[0] = ra.x; rw[1] = ra.y; rw[2] = rb.x; rw[3] = rb.y;
        }
        const uint64_t wi = gstart >> 5;
        const ulonglong2* t2 = reinterpret_cast<const ulonglong2*>(a.tg_words + (wi & ~1ull));
        const ulonglong2 ta = __ldg(t2), tb = __ldg(t2 + 1), tc = __ldg(t2 + 2);
        const bool odd = wi & 1ull;
        t[0] = odd ? ta.y : ta.x;
        t[1] = odd ? tb.x : ta.y;
        t[2] = odd ? tb.y : tb.x;
        t[3] = odd ? tc.x : tb.y;
        t[4] = odd ? tc.y : tc.x;
      } else {
#pragma unroll
        for (int w = 0; w < 4; w++) rw[w] = w < nwords ? __ldg(row + w) : 0ull;
#pragma unroll
        for (int w = 0; w < 5; w++) t[w] = w <= nwords ? __ldg(tw + w) : 0ull;
      }
#pragma unroll
      for (int w = 0; w < 4; w++) {
        if (w < nwords) {
          const uint64_t x = rw[w] ^ funnel64(t[w], t[w + 1], sh);
          uint64_t m = (x | (x >> 1)) & kEvenBits;
          if (w == nwords - 1) m &= low_bases_mask(L - 32 * w);  // row words are zero past L, the target is not
          nx += __popcll(m);
        }
      }
      if (nx > budget) return false;
    } else {
      uint64_t t0 = __ldg(tw);
      for (int w = 0; w < nwords; w++) {
        const uint64_t t1 = __ldg(tw + w + 1);
        const uint64_t x = __ldg(row + w) ^ funnel64(t0, t1, sh);
        t0 = t1;
        uint64_t m = (x | (x >> 1)) & kEvenBits;
        if (anyx) {
          const uint64_t xa = rx ? __ldg(xrow + w) : 0ull;
          const uint64_t xb = tx ? funnel64(__ldg(txw + w), __ldg(txw + w + 1), sh) & kEvenBits : 0ull;
          m = (m & ~(xa | xb)) | (xa ^ xb);  // X==X matches, X vs base mismatches (cdiff compares bytes)
        }
        if (w == nwords - 1) m &= low_bases_mask(L - 32 * w);
        nx += __popcll(m);
        if (nx > budget) return false;
      }
    }
  }

  // The pair passes through window k.
  const uint32_t slot = cj.w;
  atomicAdd(a.pass_small + ((slot * 0x9E3779B1u) >> (32 - a.lg_small)), 1u);
  if (a.pass_cnt) atomicAdd(a.pass_cnt + slot, 1u);
  n_pass++;
  if (MODE == 2 && a.slot_over[slot]) {
    const unsigned long long at = warp_agg_inc(a.n_over_inst);
    if (at < a.over_cap) a.over[at] = make_uint4(r, cj.z, (uint32_t)pos, (uint32_t)nx | ((uint32_t)k << 16));
    return false;
  }

  // Cross-window de-duplication: emit only through the lowest window index that delivers
  // this (read, gene, pos).  Window k' delivers it iff it is valid for the read, its k-mer
  // matches exactly, and -- when it would sit at target position 0 -- the literal-100 rule holds.
  // (the valid-window mask of the read: from the spare bits of its row's last word -- the line is in the L1 -- when
  // the row has them, WinCfg::vm_in_row; else from the per-read array)
  uint32_t vm = 0u;
  if (k) vm = (cfg.vm_in_row ? (uint32_t)(__ldg(row + (cfg.S - 1)) >> (64 - cfg.nwin)) : __ldg(a.validmask + r)) & ((1u << k) - 1u);
  while (vm) {
    const int k2 = __ffs(vm) - 1;
    vm &= vm - 1;
    const int q1b = cfg.windows[k2];
    if (pos + q1b == 0 && (int64_t)L > lim0) continue;
    if (!window_match(row, xrow, rx, a.tg_words, a.tg_x, tx, (uint64_t)q1b, gstart + (uint64_t)q1b, W)) continue;
    if (MODE == 2) {
      // a window whose key group is subject to truncation does not "own" the pair
      const int64_t s2 = table_find(a.tab, a.n_buckets, read_window_fp(row, xrow, rx, (uint64_t)q1b, W));
      if (s2 >= 0 && a.slot_over[s2]) continue;
    }
    return false;  // an earlier window owns this pair
  }

  atomicMin(a.best + r, (uint32_t)nx);
  rec = make_uint4(r, cj.z, (uint32_t)pos, (uint32_t)nx);
  return true;
}

// Persistent grid: blocks stride over chunks of kPairsPerThread * 256 consecutive pairs; the pair
// count lives on the device, so the launch configuration never depends on a host round trip.
// A thread owns kPairsPerThread consecutive pairs: one binary search (narrowed by block_first)
// for the first, a linear advance for the rest.  Each WARP stages its output records in its own
// shared-memory slice (positions from a ballot prefix, no atomics) and appends them with one
// global atomic per warp and chunk; there is no block-wide barrier, so a warp never waits for
// the slowest pair of another warp.  The pass counter costs one atomic per warp.
constexpr int kPairsPerThread = 4;
constexpr int kChunkPairs = kPairsPerThread * 256;
constexpr int kOutStage = 128;  // per-warp output staging (records); flushed when nearly full

#ifndef MSC_CONFIRM_CTAS
#define MSC_CONFIRM_CTAS 5
#endif
template <int MODE>
__global__ void __launch_bounds__(256, MSC_CONFIRM_CTAS) confirm_pairs_kernel(const WinCfg cfg, const ConfirmArgs a) {
  pdl_enter();
  __shared__ uint4 s_out[8][kOutStage];
  const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
  uint4* my_out = s_out[wid];
  const uint64_t n_pairs = *a.n_pairs_ptr;
  const uint64_t n_blocks = min((uint64_t)((n_pairs + kPairBlock - 1) >> kPairBlockShift), a.block_cap);
  const uint64_t n_chunks = ((n_blocks << kPairBlockShift) + kChunkPairs - 1) / kChunkPairs;
  uint32_t n_pass = 0;
  uint32_t n_out = 0;  // records staged by this warp (warp-uniform), carried across chunks
  const bool targets_have_x = *a.targets_have_x != 0ull;
  auto flush_out = [&]() {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(a.n_match, (unsigned long long)n_out);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (uint32_t t = lane; t < n_out; t += 32)
      if (base + t < a.match_cap) a.matches[base + t] = my_out[t];
    __syncwarp();
    n_out = 0;
  };
  for (uint64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const uint64_t i0 = ch * (uint64_t)kChunkPairs + (uint64_t)threadIdx.x * kPairsPerThread;
    const bool live = i0 < n_pairs && (i0 >> kPairBlockShift) < n_blocks;
    uint64_t c = 0, c_start = 0, c_end = 0;
    if (live) {
      const uint64_t b = i0 >> kPairBlockShift;
      const uint64_t clo = __ldg(a.block_first + b), chi = __ldg(a.block_first + b + 1);
      c = upper_bound_dev<uint64_t>(a.pstart, clo, chi + 1, i0) - 1;
      c_start = __ldg(a.pstart + c);
      c_end = __ldg(a.pstart + c + 1);
    }
#pragma unroll 1
    for (int j = 0; j < kPairsPerThread; j++) {
      const uint64_t i = i0 + j;
      bool has = false;
      uint4 rec = make_uint4(0u, 0u, 0u, 0u);
      if (live && i < n_pairs) {
        while (c_end <= i) {  // next candidate with at least one pair
          c++;
          c_start = c_end;
          c_end = __ldg(a.pstart + c + 1);
        }
        has = confirm_one_pair<MODE>(cfg, a, i, c, c_start, targets_have_x, rec, n_pass);
      }
      const unsigned m = __ballot_sync(0xffffffffu, has);
      if (has) my_out[n_out + __popc(m & ((1u << lane) - 1u))] = rec;
      n_out += __popc(m);
      __syncwarp();
      if (n_out > kOutStage - 32) flush_out();
    }
  }
  if (n_out) flush_out();
  n_pass = __reduce_add_sync(0xffffffffu, n_pass);
  if (lane == 0 && n_pass) atomicAdd(a.n_pass, (unsigned long long)n_pass);
}

}  // namespace msc
