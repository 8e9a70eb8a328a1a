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
  if (xm == 0) {
    // no X in the window: the adjacent pair (c_i, c_{i+1}) is the nibble of `key` at bit 2i
    uint32_t seen = 0;
    for (int i = 0; i + 1 < W; i++) seen |= 1u << (unsigned)((key >> (2 * i)) & 15ull);
    return __popc(seen);
  }
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
  uint64_t n_items;      // n_reads * nwin
  const uint64_t* fps;   // per item: key fingerprint, 0 = window not valid for the read
  uint64_t* tab_fp;
  uint32_t* tab_item0;   // first (read, window) item of the slot's key group
  uint32_t* tab_cnt;     // number of FURTHER items of the group (they go to the CSR `items`)
  int lg_slots;
  uint32_t* dup_slot;    // per item: 1 + slot if the item is a further member of its group, else 0
};
// (The number of further members, n_dup, is the grand total of the tab_cnt scan and the number of
// distinct fingerprints is n_keys - n_dup: no per-warp counter atomics in the insert kernel --
// ~10^6 same-address atomics cost more than the inserts themselves.)

// Pass A1: per read, which windows are valid (length rule + entropy rule) and the fingerprint of
// each valid window key; the key's bits are set in the Bloom front here, where the key itself
// (needed for the minimiser addressing, common.cuh) is still at hand.  item = read * nwin + window.
__global__ void __launch_bounds__(256) window_keys_kernel(const WinCfg cfg, const uint64_t* __restrict__ rd_words,
                                                          const uint64_t* __restrict__ rd_x,
                                                          const uint32_t* __restrict__ len_flags, uint64_t n_reads,
                                                          uint32_t* __restrict__ validmask, uint64_t* __restrict__ fps,
                                                          unsigned long long* __restrict__ n_keys,
                                                          unsigned long long* __restrict__ bloom, const BloomGeom geom,
                                                          const int32_t* __restrict__ nmiss, uint2* __restrict__ rmeta) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t nk = 0;
  if (r < n_reads) {
    const uint32_t lf = len_flags[r];
    const int L = (int)(lf & 0x7fffffffu);
    const bool hasx = lf >> 31;
    const uint64_t* row = rd_words + r * (uint64_t)cfg.S;
    const uint64_t* xrow = rd_x + r * (uint64_t)cfg.S;
    const uint64_t kmask = low_bases_mask(cfg.W);
    uint32_t vm = 0;
    for (int k = 0; k < cfg.nwin; k++) {
      const int q1 = cfg.windows[k], q2 = q1 + cfg.W;
      uint64_t fp = 0;
      if (L >= q2) {  // cmd/muscato_window_reads/main.go:109-112, cmd/muscato_screen/main.go:177-179
        const uint64_t key = extract32(row, (uint64_t)q1) & kmask;
        const uint64_t xm = hasx ? (extract32(xrow, (uint64_t)q1) & kmask) : 0ull;
        if (cfg.min_dinuc <= 0 || dinuc_count(key, xm, cfg.W) >= cfg.min_dinuc) {  // :183-185 / :116-118
          fp = key_fp(key, xm);
          vm |= 1u << k;
          nk++;
          uint64_t widx;
          uint32_t mlo, mhi;
          bloom_locate(key, xm, fp, cfg.W, geom, widx, mlo, mhi);
          atomicOr(bloom + widx, (unsigned long long)mlo | ((unsigned long long)mhi << 32));
        }
      }
      fps[r * (uint64_t)cfg.nwin + (uint64_t)k] = fp;
    }
    validmask[r] = vm;
    // what the confirm kernel needs of a read in one 8-byte load: length (11 bits), mismatch
    // budget nmiss(L) (11 bits, host-computed float64 table), has-X flag, valid-window mask
    rmeta[r] = make_uint2((uint32_t)L | ((uint32_t)__ldg(nmiss + L) << 11) | (lf & 0x80000000u), vm);
  }
  // one counter atomic per block (same-address atomics serialise)
  __shared__ uint32_t s_nk[8];
  nk = __reduce_add_sync(0xffffffffu, nk);
  if ((threadIdx.x & 31u) == 0) s_nk[threadIdx.x >> 5] = nk;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < 8; w++) t += s_nk[w];
    if (t) atomicAdd(n_keys, (unsigned long long)t);
  }
}

// Pass A2: one thread per item.  Claim / find the table slot of the fingerprint.  The first item of a key group lives in the slot itself (most groups have exactly
// one member); further members are flagged in dup_slot and scattered into the slot's CSR range
// by pass B.
__global__ void __launch_bounds__(256) build_insert_kernel(const BuildArgs a) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < a.n_items) {
    const uint64_t fp = __ldg(a.fps + idx);
    uint32_t dup = 0;
    if (fp) {
      const uint64_t smask = (1ull << a.lg_slots) - 1ull;
      uint64_t s = table_home(fp, a.lg_slots);
      bool first = false;
      while (true) {
        const unsigned long long cur =
            atomicCAS(reinterpret_cast<unsigned long long*>(a.tab_fp + s), 0ull, (unsigned long long)fp);
        if (cur == 0ull) { first = true; break; }
        if (cur == fp) break;
        s = (s + 1) & smask;
      }
      if (first) {
        a.tab_item0[s] = (uint32_t)idx;
      } else {
        atomicAdd(a.tab_cnt + s, 1u);
        dup = (uint32_t)s + 1u;
      }
    }
    a.dup_slot[idx] = dup;
  }
}

// Pass B1: CSR ranges for the further members.  Only slots that own further members need one, so
// instead of scanning the counts of ALL slots the first further member of a slot to arrive
// reserves the slot's range from a bump counter (tab_cnt is final by now).  tab_fill[s] ends up
// equal to tab_cnt[s] and is consumed by pass B2.
__global__ void __launch_bounds__(256) build_alloc_kernel(const uint32_t* __restrict__ dup_slot, uint64_t n_items,
                                                          const uint32_t* __restrict__ tab_cnt,
                                                          uint32_t* __restrict__ tab_fill,
                                                          uint32_t* __restrict__ tab_start,
                                                          unsigned long long* __restrict__ n_dup) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_items) return;
  const uint32_t d = __ldg(dup_slot + idx);
  if (d && atomicAdd(tab_fill + (d - 1), 1u) == 0u)
    tab_start[d - 1] = (uint32_t)atomicAdd(n_dup, (unsigned long long)__ldg(tab_cnt + (d - 1)));
}

// Pass B2: scatter the further members into their slot's CSR range.
__global__ void __launch_bounds__(256) build_fill_kernel(const uint32_t* __restrict__ dup_slot, uint64_t n_items,
                                                         const uint32_t* __restrict__ tab_start,
                                                         uint32_t* __restrict__ tab_fill,
                                                         uint32_t* __restrict__ items) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_items) return;
  const uint32_t d = __ldg(dup_slot + idx);
  if (d) items[tab_start[d - 1] + (atomicSub(tab_fill + (d - 1), 1u) - 1u)] = (uint32_t)idx;
}

// Member j of the key group in `slot` (j = 0 is stored in the slot, the rest in the CSR).
__device__ __forceinline__ uint32_t group_item(const uint32_t* __restrict__ tab_item0,
                                               const uint32_t* __restrict__ tab_start,
                                               const uint32_t* __restrict__ items, uint32_t slot, uint32_t j) {
  return j == 0 ? __ldg(tab_item0 + slot) : __ldg(items + __ldg(tab_start + slot) + (j - 1));
}

}  // namespace msc
