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

// The same for a wide window (32 < W <= 64) held in two words.
__device__ __forceinline__ int dinuc_count_wide(uint64_t k0, uint64_t k1, uint64_t x0, uint64_t x1, int W) {
  auto sym = [&](int i) -> uint32_t {
    const uint64_t k = i < 32 ? k0 : k1, x = i < 32 ? x0 : x1;
    const int s = 2 * (i & 31);
    return ((x >> s) & 1ull) ? 4u : (uint32_t)((k >> s) & 3ull);
  };
  uint32_t seen = 0;
  uint32_t prev = sym(0);
  for (int i = 1; i < W; i++) {
    const uint32_t cur = sym(i);
    seen |= 1u << (5u * prev + cur);
    prev = cur;
  }
  return __popc(seen);
}

struct BuildArgs {
  // reads
  const uint64_t* rd_words;
  const uint64_t* rd_x;
  const uint32_t* len_flags;
  uint64_t n_reads;
  const int32_t* nmiss;  // [MRL + 1], host-computed float64 table (cmd/muscato_confirm/main.go:198)
  // per read outputs
  uint32_t* validmask;
  uint2* rmeta;
  unsigned long long* n_keys;
  // key table
  uint64_t* tab_fp;
  uint32_t* tab_item0;   // first (read, window) item of the slot's key group
  uint32_t* tab_cnt;     // number of FURTHER items of the group (they go to the CSR `items`)
  int lg_slots;
  uint32_t* dup_slot;    // per item: 1 + slot if the item is a further member of its group, else 0
  // Bloom front
  unsigned long long* bloom;
  BloomGeom geom;
};
// (The number of further members, n_dup, is the grand total of the bump allocation in pass B1 and
// the number of distinct fingerprints is n_keys - n_dup: no per-warp counter atomics in the insert
// kernel -- ~10^6 same-address atomics cost more than the inserts themselves.)

#ifndef MSC_INSERT_BATCH
#define MSC_INSERT_BATCH 1
#endif
#ifndef MSC_INSERT_CTAS
#define MSC_INSERT_CTAS 8
#endif
constexpr int kInsertBatch = MSC_INSERT_BATCH;  // home buckets in flight per thread

// Pass A: one thread per read.  Which windows are valid (length rule + entropy rule), the
// fingerprint of each valid window key, its Bloom bits, and the claim of its table slot: the
// first item of a key group lives in the slot itself (most groups have exactly one member);
// further members are flagged in dup_slot and scattered into the slot's CSR range by pass B.
// The home buckets (32 bytes, one 256-bit load each) of up to kInsertBatch windows are fetched
// before any claim is made; a claim is then one atomicCAS on the first slot seen free.
// item = read * nwin + window.
__global__ void __launch_bounds__(256, MSC_INSERT_CTAS) build_keys_insert_kernel(const WinCfg cfg, const BuildArgs a) {
  pdl_enter();
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t nk = 0;
  if (r < a.n_reads) {
    const uint32_t lf = a.len_flags[r];
    const int L = (int)(lf & 0x7fffffffu);
    const bool hasx = lf >> 31;
    const uint64_t* row = a.rd_words + r * (uint64_t)cfg.S;
    const uint64_t* xrow = a.rd_x + r * (uint64_t)cfg.S;
    const bool wide = cfg.W > 32;
    const uint64_t kmask = low_bases_mask(min(cfg.W, 32));
    const uint64_t kmask1 = wide ? low_bases_mask(cfg.W - 32) : 0ull;
    const uint64_t bmask = (1ull << (a.lg_slots - 2)) - 1ull;
    uint32_t vm = 0;
    for (int k0 = 0; k0 < cfg.nwin; k0 += kInsertBatch) {
      uint64_t fp[kInsertBatch], bk[kInsertBatch];
      uint64_t q[kInsertBatch][4];
#pragma unroll
      for (int u = 0; u < kInsertBatch; u++) {
        const int k = k0 + u;
        fp[u] = 0;
        bk[u] = 0;
        q[u][0] = q[u][1] = q[u][2] = q[u][3] = 0;
        if (k < cfg.nwin) {
          const int q1 = cfg.windows[k], q2 = q1 + cfg.W;
          if (L >= q2) {  // cmd/muscato_window_reads/main.go:109-112, cmd/muscato_screen/main.go:177-179
            const uint64_t key = extract32(row, (uint64_t)q1) & kmask;
            const uint64_t xm0 = hasx ? (extract32(xrow, (uint64_t)q1) & kmask) : 0ull;
            uint64_t key1 = 0, xm1 = 0;
            if (wide) {  // bases 32..W-1 of the window
              key1 = extract32(row, (uint64_t)q1 + 32) & kmask1;
              xm1 = hasx ? (extract32(xrow, (uint64_t)q1 + 32) & kmask1) : 0ull;
            }
            const uint64_t xm = xm0 | xm1;
            const int nd = cfg.min_dinuc <= 0 ? 0
                           : wide ? dinuc_count_wide(key, key1, xm0, xm1, cfg.W) : dinuc_count(key, xm0, cfg.W);
            if (cfg.min_dinuc <= 0 || nd >= cfg.min_dinuc) {  // :183-185 / :116-118
              fp[u] = wide ? key_fp_wide(key, key1, xm0, xm1) : key_fp(key, xm0);
              vm |= 1u << k;
              nk++;
              uint64_t widx;
              uint32_t mlo, mhi;
              bloom_locate(key, xm, fp[u], cfg.W, a.geom, widx, mlo, mhi, key1);
              atomicOr(a.bloom + widx, (unsigned long long)mlo | ((unsigned long long)mhi << 32));
              bk[u] = table_home_bucket(fp[u], a.lg_slots);
              ldcg256(a.tab_fp + (bk[u] << 2), q[u][0], q[u][1], q[u][2], q[u][3]);  // the home bucket as it stands
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kInsertBatch; u++) {
        const int k = k0 + u;
        if (k >= cfg.nwin) break;
        const uint64_t item = r * (uint64_t)cfg.nwin + (uint64_t)k;
        uint32_t dup = 0;
        if (fp[u]) {
          // first slot of the probe sequence that holds fp (further member) or that this thread
          // claims with a CAS (first member); a slot seen free may have been taken meanwhile --
          // the CAS returns what is there now
          int64_t slot = -1;
          bool first = false;
          while (slot < 0) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
              if (slot >= 0) break;
              uint64_t v = q[u][i];
              if (v == 0ull)
                v = atomicCAS(reinterpret_cast<unsigned long long*>(a.tab_fp + (bk[u] << 2) + i), 0ull, (unsigned long long)fp[u]);
              if (v == 0ull) {
                slot = (int64_t)((bk[u] << 2) + i);
                first = true;
              } else if (v == fp[u]) {
                slot = (int64_t)((bk[u] << 2) + i);
              }
            }
            if (slot < 0) {  // bucket full of other keys: next bucket
              bk[u] = (bk[u] + 1) & bmask;
              ldcg256(a.tab_fp + (bk[u] << 2), q[u][0], q[u][1], q[u][2], q[u][3]);
            }
          }
          if (first) {
            a.tab_item0[slot] = (uint32_t)item;
          } else {
            atomicAdd(a.tab_cnt + slot, 1u);
            dup = (uint32_t)slot + 1u;
          }
        }
        a.dup_slot[item] = dup;
      }
    }
    a.validmask[r] = vm;
    // what the confirm kernel needs of a read in one 8-byte load: length (11 bits), mismatch
    // budget nmiss(L) (11 bits), has-X flag, valid-window mask
    a.rmeta[r] = make_uint2((uint32_t)L | ((uint32_t)__ldg(a.nmiss + L) << 11) | (lf & 0x80000000u), vm);
  }
  // one counter atomic per block (same-address atomics serialise)
  __shared__ uint32_t s_nk[8];
  nk = __reduce_add_sync(0xffffffffu, nk);
  if ((threadIdx.x & 31u) == 0) s_nk[threadIdx.x >> 5] = nk;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < 8; w++) t += s_nk[w];
    if (t) atomicAdd(a.n_keys, (unsigned long long)t);
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
  pdl_enter();
  // the reservations of a block are summed in shared memory: ONE bump of the global counter per
  // block (same-address global atomics serialise)
  __shared__ uint32_t s_total;
  __shared__ unsigned long long s_base;
  if (threadIdx.x == 0) s_total = 0u;
  __syncthreads();
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t d = idx < n_items ? __ldg(dup_slot + idx) : 0u;
  const bool first = d && atomicAdd(tab_fill + (d - 1), 1u) == 0u;
  uint32_t local = 0;
  if (first) local = atomicAdd(&s_total, __ldg(tab_cnt + (d - 1)));
  __syncthreads();
  if (threadIdx.x == 0 && s_total) s_base = atomicAdd(n_dup, (unsigned long long)s_total);
  __syncthreads();
  if (first) tab_start[d - 1] = (uint32_t)(s_base + local);
}

// Pass B2: scatter the further members into their slot's CSR range.  A CSR entry is 16 bytes:
// (item, read record) -- the confirm kernel gets the read's length / budget / window mask with
// the item itself instead of through one more dependent look-up.
__global__ void __launch_bounds__(256) build_fill_kernel(const uint32_t* __restrict__ dup_slot, uint64_t n_items,
                                                         const uint32_t* __restrict__ tab_start,
                                                         uint32_t* __restrict__ tab_fill,
                                                         const uint2* __restrict__ rmeta, uint32_t nwin,
                                                         uint4* __restrict__ items) {
  pdl_enter();
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_items) return;
  const uint32_t d = __ldg(dup_slot + idx);
  if (d) {
    const uint2 rm = __ldg(rmeta + (uint32_t)idx / nwin);
    items[tab_start[d - 1] + (atomicSub(tab_fill + (d - 1), 1u) - 1u)] = make_uint4((uint32_t)idx, rm.x, rm.y, 0u);
  }
}

}  // namespace msc
