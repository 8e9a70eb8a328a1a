// build.cuh -- kernel family (1b): window-key extraction from the packed reads and the
// device-resident open-addressing key table + blocked Bloom front.
//
// Replaces buildBloom (cmd/muscato_screen/main.go:116-207), muscato_window_reads
// (cmd/muscato_window_reads/main.go:100-140) and sortWindows (cmd/muscato/main.go:237-304):
// instead of NumHash rolling hashes into one BloomSize-bit array per window, every valid
// (read, window) key goes into ONE table keyed on the k-mer, whose slot owns a CSR range of
// (read, window) items -- the "block of lines sharing field 0" of the merge join
// (cmd/muscato_confirm/main.go:98-148).
#pragma once
#include "common.cuh"

namespace msc {

struct WinCfg {
  int nwin, W, MRL, S, min_dinuc;
  int windows[32];
};

// utils.CountDinuc (utils/entropy.go:5-40) on a packed window: number of distinct adjacent
// symbol pairs over a 5-letter alphabet (4 bases + X).  The count is invariant under
// relabelling of the symbols, so the packed codes are used directly.
__device__ __forceinline__ int dinuc_count(uint64_t key, uint64_t xm, int W) {
  uint32_t seen = 0;
  uint32_t prev = (xm & 1ull) ? 4u : (uint32_t)(key & 3ull);
  for (int i = 1; i < W; i++) {
    const uint32_t cur = ((xm >> (2 * i)) & 1ull) ? 4u : (uint32_t)((key >> (2 * i)) & 3ull);
    seen |= 1u << (5u * prev + cur);
    prev = cur;
  }
  return __popc(seen);
}

struct BuildArgs {
  const uint64_t* rd_words;
  const uint64_t* rd_x;
  const uint32_t* len_flags;
  uint64_t n_reads;
  uint64_t* tab_fp;
  uint32_t* tab_cnt;
  int lg_slots;
  unsigned long long* bloom;
  int lg_bloom;
  uint32_t* validmask;
  unsigned long long* n_keys;    // valid (read, window) keys
  unsigned long long* n_groups;  // distinct fingerprints
};

// Pass A: per read, evaluate every window (length + entropy rule), claim / find the table
// slot of its fingerprint, count it, and set the Bloom bits.
__global__ void __launch_bounds__(256) build_insert_kernel(const WinCfg cfg, const BuildArgs a) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t nk = 0, ng = 0;
  if (r < a.n_reads) {
    const uint32_t lf = a.len_flags[r];
    const int L = (int)(lf & 0x7fffffffu);
    const bool hasx = lf >> 31;
    const uint64_t* row = a.rd_words + r * (uint64_t)cfg.S;
    const uint64_t* xrow = a.rd_x + r * (uint64_t)cfg.S;
    const uint64_t kmask = low_bases_mask(cfg.W);
    const uint64_t smask = (1ull << a.lg_slots) - 1ull;
    uint32_t vm = 0;
    for (int k = 0; k < cfg.nwin; k++) {
      const int q1 = cfg.windows[k], q2 = q1 + cfg.W;
      if (L < q2) continue;  // cmd/muscato_window_reads/main.go:109-112, cmd/muscato_screen/main.go:177-179
      const uint64_t key = extract32(row, (uint64_t)q1) & kmask;
      const uint64_t xm = hasx ? (extract32(xrow, (uint64_t)q1) & kmask) : 0ull;
      if (cfg.min_dinuc > 0 && dinuc_count(key, xm, cfg.W) < cfg.min_dinuc) continue;  // :183-185 / :116-118
      vm |= 1u << k;
      nk++;
      const uint64_t fp = key_fp(key, xm);
      uint64_t s = table_home(fp, a.lg_slots);
      while (true) {
        const unsigned long long cur =
            atomicCAS(reinterpret_cast<unsigned long long*>(a.tab_fp + s), 0ull, (unsigned long long)fp);
        if (cur == 0ull) { ng++; break; }
        if (cur == fp) break;
        s = (s + 1) & smask;
      }
      atomicAdd(a.tab_cnt + s, 1u);
      const unsigned long long bm = (unsigned long long)bloom_mask_lo(fp) | ((unsigned long long)bloom_mask_hi(fp) << 32);
      atomicOr(a.bloom + bloom_index(fp, a.lg_bloom), bm);
    }
    a.validmask[r] = vm;
  }
  nk = __reduce_add_sync(0xffffffffu, nk);
  ng = __reduce_add_sync(0xffffffffu, ng);
  if ((threadIdx.x & 31u) == 0) {
    if (nk) atomicAdd(a.n_keys, (unsigned long long)nk);
    if (ng) atomicAdd(a.n_groups, (unsigned long long)ng);
  }
}

// Pass B (after the exclusive scan of tab_cnt into tab_start): scatter the (read, window)
// items into their slot's CSR range.  item = read * nwin + window.
__global__ void __launch_bounds__(256) build_fill_kernel(const WinCfg cfg, const uint64_t* __restrict__ rd_words,
                                                         const uint64_t* __restrict__ rd_x,
                                                         const uint32_t* __restrict__ len_flags,
                                                         const uint32_t* __restrict__ validmask, uint64_t n_reads,
                                                         const uint64_t* __restrict__ tab_fp,
                                                         const uint32_t* __restrict__ tab_start,
                                                         uint32_t* __restrict__ tab_fill, int lg_slots,
                                                         uint32_t* __restrict__ items) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  uint32_t vm = validmask[r];
  if (!vm) return;
  const bool hasx = len_flags[r] >> 31;
  const uint64_t* row = rd_words + r * (uint64_t)cfg.S;
  const uint64_t* xrow = rd_x + r * (uint64_t)cfg.S;
  const uint64_t kmask = low_bases_mask(cfg.W);
  while (vm) {
    const int k = __ffs(vm) - 1;
    vm &= vm - 1;
    const int q1 = cfg.windows[k];
    const uint64_t key = extract32(row, (uint64_t)q1) & kmask;
    const uint64_t xm = hasx ? (extract32(xrow, (uint64_t)q1) & kmask) : 0ull;
    const int64_t s = table_find(tab_fp, lg_slots, key_fp(key, xm));
    const uint32_t at = tab_start[s] + atomicAdd(tab_fill + s, 1u);
    items[at] = (uint32_t)(r * (uint64_t)cfg.nwin + (uint64_t)k);
  }
}

}  // namespace msc
