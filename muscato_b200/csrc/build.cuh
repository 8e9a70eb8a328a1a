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
  // layout of the 32-bit read record word that travels with every (read, window) item:
  //   bits [0, lbits) read length, [lbits, lbits + nbits) mismatch budget nmiss(L), then 2 * sk bits of SKETCH
  //   (sk read bases next to the window, see item_sketch), bit 31 = the read contains X
  int lbits, nbits, sk;
  // 1 = the read's valid-window mask also sits in the top nwin bits of the LAST word of its packed row (bases the
  // row does not use: 2 * (32 * S - MRL) >= nwin), so that confirm finds it in the row line it has just fetched
  // instead of in a second random line (measured: 10 % of the confirm kernel's stall samples, 8.5 GB at configs[2])
  int vm_in_row;
  int windows[32];
};

__host__ __device__ __forceinline__ int rmx_len(const WinCfg& c, uint32_t rmx) { return (int)(rmx & ((1u << c.lbits) - 1u)); }
__host__ __device__ __forceinline__ int rmx_budget(const WinCfg& c, uint32_t rmx) {
  return (int)((rmx >> c.lbits) & ((1u << c.nbits) - 1u));
}
__host__ __device__ __forceinline__ uint32_t rmx_sketch(const WinCfg& c, uint32_t rmx) {
  return (rmx >> (c.lbits + c.nbits)) & ((1u << (2 * c.sk)) - 1u);
}
// First read base of the sketch of window k of a read of length L: the sk bases right of the window when the read
// has them, else the sk bases left of it, else none (-1).  Build and confirm evaluate the same rule.
__host__ __device__ __forceinline__ int sketch_start(const WinCfg& c, int k, int L) {
  const int q1 = c.windows[k], q2 = q1 + c.W;
  if (c.sk == 0) return -1;
  if (q2 + c.sk <= L) return q2;
  if (q1 >= c.sk) return q1 - c.sk;
  return -1;
}

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

// ---------------------------------------------------------------------------------------------
// Key table build (round 2).  The table layout is described in common.cuh: buckets of five slots,
// one 128-byte line each, holding the fingerprints AND the per-group records {item0, rmx0, start, cnt};
// items[] = uint2 {item, rmx} of the FURTHER members of every key group (CSR).
//
// The build is PARTITIONED so that the random accesses of the insert stay inside the L2: the keys
// are first distributed into partitions by the range of their home bucket (count, offsets, scatter
// through a shared-memory stage: sequential traffic only), then inserted partition after partition
// -- the table region one partition touches is 16 MB, so the bucket reads, the CAS claims and
// the record writes are served by the L2 instead of one 128-byte HBM line each (the round-1 insert
// moved 383 B of DRAM traffic per key).
// Replaces buildBloom (cmd/muscato_screen/main.go:116-207) + sortWindows (cmd/muscato/main.go:237-304).
// ---------------------------------------------------------------------------------------------
struct BuildArgs {
  // reads
  const uint64_t* rd_words;
  const uint64_t* rd_x;
  const uint32_t* len_flags;
  uint64_t n_reads;
  uint64_t* rd_words_rw;        // = rd_words (build_windows_kernel adds the valid-window mask to the row's spare bits)
  uint64_t w_begin, w_end;      // build_windows_kernel: the reads of this launch (a chunk of a chunked upload)
  const int32_t* nmiss;  // [MRL + 1], host-computed float64 table (cmd/muscato_confirm/main.go:198)
  // per read outputs
  uint32_t* validmask;
  uint2* rmeta;
  unsigned long long* n_keys;
  // key table
  uint8_t* tab;
  TableGeom tg;
  // partitioning
  unsigned int* part_count;     // [n_parts]: keys per partition (pass A), then the write cursors (pass B)
  uint4* recs;                  // partition-ordered key records {fp.lo, fp.hi, item, rmx}
  // further members
  uint4* dups;                  // {slot, item, rmx, ordinal}
  unsigned long long* n_dup;
  unsigned long long* n_alloc;  // bump allocator of the CSR
  unsigned long long* insert_cursor;  // build_insert_kernel: next record to hand out
  uint2* items;
  // Bloom front
  unsigned long long* bloom;
  BloomGeom geom;
};

constexpr int kMaxParts = 1024;
// The per-partition counters / write cursors sit one per 128-byte line (kPartStride words apart): every block of the
// window pass and of the scatter bumps all of them, and atomics on one line serialise in its L2 slice (packed, the 1024
// counters are 32 lines).
constexpr int kPartStride = 32;

__device__ __forceinline__ uint32_t key_partition(uint64_t fp, const TableGeom& tg) {
  return (uint32_t)(table_home_bucket(fp, tg.n_buckets) >> tg.lg_bpp);
}

// Fingerprint of window k of read r, as every pass derives it; optionally the key words and X masks.
__device__ __forceinline__ uint64_t window_fp(const WinCfg& cfg, const uint64_t* __restrict__ row,
                                              const uint64_t* __restrict__ xrow, bool hasx, int k, uint64_t* key_out,
                                              uint64_t* key1_out, uint64_t* xm0_out, uint64_t* xm1_out) {
  const int q1 = cfg.windows[k];
  const bool wide = cfg.W > 32;
  const uint64_t kmask = low_bases_mask(min(cfg.W, 32));
  const uint64_t key = extract32(row, (uint64_t)q1) & kmask;
  const uint64_t xm0 = hasx ? (extract32(xrow, (uint64_t)q1) & kmask) : 0ull;
  uint64_t key1 = 0, xm1 = 0;
  if (wide) {  // bases 32..W-1 of the window
    const uint64_t kmask1 = low_bases_mask(cfg.W - 32);
    key1 = extract32(row, (uint64_t)q1 + 32) & kmask1;
    xm1 = hasx ? (extract32(xrow, (uint64_t)q1 + 32) & kmask1) : 0ull;
  }
  if (key_out) *key_out = key;
  if (key1_out) *key1_out = key1;
  if (xm0_out) *xm0_out = xm0;
  if (xm1_out) *xm1_out = xm1;
  return wide ? key_fp_wide(key, key1, xm0, xm1) : key_fp(key, xm0);
}

// Clears the fingerprints of every bucket: the first 64 of its 128 bytes (two whole sectors; the
// records behind them are only ever read after their fingerprint was claimed).  Four threads per bucket.
__global__ void __launch_bounds__(256) table_clear_kernel(uint8_t* __restrict__ tab, uint64_t n_buckets) {
  pdl_enter();
  const uint64_t n = n_buckets * 4;
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    *reinterpret_cast<uint4*>(tab + (i >> 2) * (uint64_t)kBucketBytes + (i & 3ull) * 16ull) = z;
}

// Pass A: one thread per read (grid-stride).  Which windows are valid (length rule + entropy rule,
// cmd/muscato_window_reads/main.go:109-118, cmd/muscato_screen/main.go:174-185), the read record,
// the Bloom bits of every valid key, and the number of keys per partition (shared-memory histogram,
// one global atomic per block and non-empty partition).  item = read * nwin + window.
__global__ void __launch_bounds__(256) build_windows_kernel(const WinCfg cfg, const BuildArgs a) {
  pdl_enter();
  __shared__ unsigned int s_hist[kMaxParts];
  const int P = (int)a.tg.n_parts;
  for (int p = threadIdx.x; p < P; p += blockDim.x) s_hist[p] = 0u;
  __syncthreads();
  uint32_t nk = 0;
  const uint64_t pol_keep = l2_policy_evict_last();
  for (uint64_t r = a.w_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < a.w_end; r += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t lf = a.len_flags[r];
    const int L = (int)(lf & 0x7fffffffu);
    const bool hasx = lf >> 31;
    const uint64_t* row = a.rd_words + r * (uint64_t)cfg.S;
    const uint64_t* xrow = a.rd_x + r * (uint64_t)cfg.S;
    uint32_t vm = 0;
    for (int k = 0; k < cfg.nwin; k++) {
      if (L < cfg.windows[k] + cfg.W) continue;  // cmd/muscato_window_reads/main.go:109-112, cmd/muscato_screen/main.go:177-179
      uint64_t key, key1, xm0, xm1;
      const uint64_t fp = window_fp(cfg, row, xrow, hasx, k, &key, &key1, &xm0, &xm1);
      if (cfg.min_dinuc > 0) {  // :183-185 / :116-118
        const int nd = cfg.W > 32 ? dinuc_count_wide(key, key1, xm0, xm1, cfg.W) : dinuc_count(key, xm0, cfg.W);
        if (nd < cfg.min_dinuc) continue;
      }
      const uint64_t xm = xm0 | xm1;
      vm |= 1u << k;
      nk++;
      uint64_t widx;
      uint32_t mlo, mhi;
      bloom_locate(key, xm, fp, cfg.W, a.geom, widx, mlo, mhi, key1);
      // the front's words are the one structure this pass re-uses: last out of the L2 (every RED that misses costs a
      // sector read and a sector write in HBM: ncu, 1.6 sectors of each per key with a 128 MB map and default priority)
      red_or64_hint(a.bloom + widx, (unsigned long long)mlo | ((unsigned long long)mhi << 32), pol_keep);
      atomicAdd(&s_hist[key_partition(fp, a.tg)], 1u);
    }
    __stcs(a.validmask + r, vm);  // (streaming stores: written once, read by later kernels from HBM anyway)
    if (cfg.vm_in_row && vm) a.rd_words_rw[r * (uint64_t)cfg.S + (uint64_t)(cfg.S - 1)] |= (uint64_t)vm << (64 - cfg.nwin);
    // what the confirm kernel needs of a read: length, mismatch budget nmiss(L), has-X flag (WinCfg);
    // .y = valid-window mask
    __stcs(a.rmeta + r, make_uint2((uint32_t)L | ((uint32_t)__ldg(a.nmiss + L) << cfg.lbits) | (lf & 0x80000000u), vm));
  }
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const unsigned int c = s_hist[p];
    if (c) atomicAdd(a.part_count + (size_t)p * kPartStride, c);
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

// Partition sizes -> start offsets (in place: part_count becomes the write cursor of every partition).
// One block; P <= kMaxParts.
__global__ void __launch_bounds__(kMaxParts) build_offsets_kernel(unsigned int* __restrict__ part_count, int P) {
  pdl_enter();
  __shared__ unsigned int s[kMaxParts];
  const int t = threadIdx.x;
  const unsigned int v = t < P ? part_count[(size_t)t * kPartStride] : 0u;
  s[t] = v;
  __syncthreads();
  for (int o = 1; o < kMaxParts; o <<= 1) {
    const unsigned int add = t >= o ? s[t - o] : 0u;
    __syncthreads();
    s[t] += add;
    __syncthreads();
  }
  if (t < P) part_count[(size_t)t * kPartStride] = s[t] - v;
}

// Pass B: distribute the key records into partition order.  One thread per (read, window) item; a
// block stages kStageKeys items in shared memory (dynamic, 98 KB: two blocks per SM), reserves one run per
// non-empty partition with a single global atomic each and writes the runs of 16-byte records.  The
// stage is as large as two resident blocks allow: the global atomics on the 1024 partition cursors
// (one per partition and flush) were what bounded the 2048-record version (1.2 TB/s).
constexpr int kScatterThreads = 512;
constexpr int kStageRounds = 10;
constexpr int kStageKeys = kStageRounds * kScatterThreads;  // 5120 records per flush: ~5 per partition and global atomic at 1024 partitions

struct ScatterSmem {
  alignas(16) uint4 rec[kStageKeys];
  unsigned int cnt[kMaxParts];
  unsigned int base[kMaxParts];
  uint16_t part[kStageKeys];
};

__global__ void __launch_bounds__(kScatterThreads, 2) build_scatter_kernel(const WinCfg cfg, const BuildArgs a) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char scatter_smem[];
  ScatterSmem& sm = *reinterpret_cast<ScatterSmem*>(scatter_smem);
  uint4* s_rec = sm.rec;
  uint16_t* s_part = sm.part;
  unsigned int* s_cnt = sm.cnt;
  unsigned int* s_base = sm.base;
  const int P = (int)a.tg.n_parts;
  const uint64_t n_items = a.n_reads * (uint64_t)cfg.nwin;
  const uint64_t n_tiles = (n_items + kStageKeys - 1) / kStageKeys;
  for (int p = threadIdx.x; p < P; p += kScatterThreads) s_cnt[p] = 0u;
  __syncthreads();
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll 1
    for (int j = 0; j < kStageRounds; j++) {
      const int idx = j * kScatterThreads + (int)threadIdx.x;
      const uint64_t item = tile * (uint64_t)kStageKeys + (uint64_t)idx;
      uint32_t part = 0xFFFFu;
      if (item < n_items) {
        const uint64_t r = cfg.nwin == 1 ? item : (uint64_t)((uint32_t)item / (uint32_t)cfg.nwin);  // items are 32-bit
        const int k = (int)(item - r * (uint64_t)cfg.nwin);
        if ((__ldg(a.validmask + r) >> k) & 1u) {
          const uint32_t rmx = __ldg(&a.rmeta[r].x);
          const uint64_t fp = window_fp(cfg, a.rd_words + r * (uint64_t)cfg.S, a.rd_x + r * (uint64_t)cfg.S, rmx >> 31, k,
                                        nullptr, nullptr, nullptr, nullptr);
          part = key_partition(fp, a.tg);
          // the item's sketch rides in the spare bits of the read record word (X-free reads only)
          uint32_t rw = rmx;
          const int s0 = (rmx >> 31) ? -1 : sketch_start(cfg, k, rmx_len(cfg, rmx));
          if (s0 >= 0)
            rw |= (uint32_t)(extract32(a.rd_words + r * (uint64_t)cfg.S, (uint64_t)s0) & low_bases_mask(cfg.sk)) << (cfg.lbits + cfg.nbits);
          s_rec[idx] = make_uint4((uint32_t)fp, (uint32_t)(fp >> 32), (uint32_t)item, rw);
          atomicAdd(&s_cnt[part], 1u);
        }
      }
      s_part[idx] = (uint16_t)part;
    }
    __syncthreads();
    for (int p = threadIdx.x; p < P; p += kScatterThreads) {
      const unsigned int c = s_cnt[p];
      if (c) s_base[p] = atomicAdd(a.part_count + (size_t)p * kPartStride, c);
      s_cnt[p] = 0u;
    }
    __syncthreads();
#pragma unroll 1
    for (int j = 0; j < kStageRounds; j++) {
      const int idx = j * kScatterThreads + (int)threadIdx.x;
      const uint32_t part = s_part[idx];
      if (part != 0xFFFFu) a.recs[s_base[part] + atomicAdd(&s_cnt[part], 1u)] = s_rec[idx];
    }
    __syncthreads();
    for (int p = threadIdx.x; p < P; p += kScatterThreads) s_cnt[p] = 0u;
    __syncthreads();
  }
}

// Pass C: insert the records in partition order (one thread per record).  The first member of a key
// group claims a slot with one CAS on the first slot it saw free in the bucket (the CAS returns what
// is there now if another thread was faster) and writes the slot's record; further members are
// appended to the dups list (their CSR position is settled by passes D-F, after every claim is done).
//
// The records are handed out in ORDER, one block-sized chunk per grab of a global cursor (insert_cursor, zero-filled with
// the other counters), so that the whole grid works on one front of ~3e5 consecutive records = one partition = 16 MB of
// table at any time: with a static grid-stride split the warps drift several partitions apart and the bucket lines of a
// partition left the L2 between the ~2 keys that touch each of them (ncu: one line read from HBM per KEY, not per bucket).
constexpr int kInsPerThread = 1;  // records per thread and round (2 was measured slower: 11.3 vs 10.5 ms at configs[2])

__global__ void __launch_bounds__(256) build_insert_kernel(const BuildArgs a) {
  pdl_enter();
  __shared__ unsigned long long s_chunk, s_dup_base;
  __shared__ uint32_t s_wcnt[8];
  const uint64_t n = *a.n_keys;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  while (true) {
    if (threadIdx.x == 0) s_chunk = atomicAdd(a.insert_cursor, (unsigned long long)(blockDim.x * kInsPerThread));
    __syncthreads();
    if (s_chunk >= n) break;  // block-uniform
    uint4 rc[kInsPerThread];
    uint64_t fp[kInsPerThread], bk[kInsPerThread];
    int64_t slot[kInsPerThread];
    bool first[kInsPerThread], valid[kInsPerThread];
    uint64_t q[kInsPerThread][kBucketSlots];
#pragma unroll
    for (int u = 0; u < kInsPerThread; u++) {
      const uint64_t i = s_chunk + (uint64_t)u * blockDim.x + threadIdx.x;
      valid[u] = i < n;
      rc[u] = valid[u] ? a.recs[i] : make_uint4(0u, 0u, 0u, 0u);
      fp[u] = (uint64_t)rc[u].x | ((uint64_t)rc[u].y << 32);
      bk[u] = table_home_bucket(fp[u], a.tg.n_buckets);
      slot[u] = -1;
      first[u] = false;
      if (valid[u]) {
        // the bucket as it stands (L2: other SMs are claiming slots); both records' buckets are requested before either is used
        ldcg256(bucket_ptr(a.tab, bk[u]), q[u][0], q[u][1], q[u][2], q[u][3]);
        q[u][4] = ldcg64(bucket_ptr(a.tab, bk[u]) + 32);
      }
    }
#pragma unroll
    for (int u = 0; u < kInsPerThread; u++) {
      if (!valid[u]) continue;
      while (true) {
        uint8_t* bp = bucket_ptr(a.tab, bk[u]);
#pragma unroll
        for (int s = 0; s < kBucketSlots; s++) {
          if (slot[u] >= 0) break;
          uint64_t v = q[u][s];
          if (v == 0ull) v = atomicCAS(reinterpret_cast<unsigned long long*>(bp) + s, 0ull, (unsigned long long)fp[u]);
          if (v == 0ull) {
            slot[u] = (int64_t)(bk[u] * kBucketSlots + s);
            first[u] = true;
          } else if (v == fp[u]) {
            slot[u] = (int64_t)(bk[u] * kBucketSlots + s);
          }
        }
        if (slot[u] >= 0) break;
        bk[u] = bk[u] + 1 == a.tg.n_buckets ? 0 : bk[u] + 1;  // bucket full of other keys: next bucket
        ldcg256(bucket_ptr(a.tab, bk[u]), q[u][0], q[u][1], q[u][2], q[u][3]);
        q[u][4] = ldcg64(bucket_ptr(a.tab, bk[u]) + 32);
      }
      if (first[u]) *slot_rec_ptr(a.tab, (uint64_t)slot[u]) = make_uint4(rc[u].z, rc[u].w, 0u, 0u);
    }
    // further members: ONE bump of the global list cursor per block and round (a returning atomic per warp on one
    // address -- 9.4e6 of them at configs[2] -- was what bounded this kernel, not its HBM traffic)
    unsigned bal[kInsPerThread];
    uint32_t mine = 0;
#pragma unroll
    for (int u = 0; u < kInsPerThread; u++) {
      bal[u] = __ballot_sync(0xffffffffu, valid[u] && !first[u]);
      mine += __popc(bal[u]);
    }
    if (lane == 0) s_wcnt[warp] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t t = 0;
      for (int w = 0; w < 8; w++) {
        const uint32_t c = s_wcnt[w];
        s_wcnt[w] = t;
        t += c;
      }
      s_dup_base = t ? atomicAdd(a.n_dup, (unsigned long long)t) : 0ull;
    }
    __syncthreads();
    {
      unsigned long long at = s_dup_base + s_wcnt[warp];
#pragma unroll
      for (int u = 0; u < kInsPerThread; u++) {
        if (valid[u] && !first[u]) a.dups[at + __popc(bal[u] & ((1u << lane) - 1u))] = make_uint4((uint32_t)slot[u], rc[u].z, rc[u].w, 0u);
        at += __popc(bal[u]);
      }
    }
    __syncthreads();  // s_chunk, s_wcnt, s_dup_base are rewritten by the next round
  }
}

// Pass D: every further member takes its ordinal inside its group (the cnt word of the slot record counts them).
__global__ void __launch_bounds__(256) build_dup_count_kernel(const BuildArgs a) {
  pdl_enter();
  const uint64_t n = *a.n_dup;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t slot = a.dups[i].x;
    a.dups[i].w = atomicAdd(&slot_rec_ptr(a.tab, slot)->w, 1u);
  }
}

// Pass E: the member with ordinal 0 reserves the group's CSR range (counts are final by now); the
// reservations of a block are summed in shared memory: ONE bump of the global counter per block.
__global__ void __launch_bounds__(256) build_dup_alloc_kernel(const BuildArgs a) {
  pdl_enter();
  __shared__ uint32_t s_total;
  __shared__ unsigned long long s_base;
  const uint64_t n = *a.n_dup;
  const uint64_t n_rounds = (n + (uint64_t)gridDim.x * blockDim.x - 1) / ((uint64_t)gridDim.x * blockDim.x);
  for (uint64_t rd = 0; rd < n_rounds; rd++) {
    const uint64_t i = (rd * gridDim.x + blockIdx.x) * (uint64_t)blockDim.x + threadIdx.x;
    if (threadIdx.x == 0) s_total = 0u;
    __syncthreads();
    uint32_t slot = 0, local = 0;
    bool head = false;
    if (i < n) {
      const uint4 d = a.dups[i];
      if (d.w == 0u) {
        head = true;
        slot = d.x;
        local = atomicAdd(&s_total, slot_rec_ptr(a.tab, slot)->w);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_total) s_base = atomicAdd(a.n_alloc, (unsigned long long)s_total);
    __syncthreads();
    if (head) slot_rec_ptr(a.tab, slot)->z = (uint32_t)(s_base + local);
    __syncthreads();
  }
}

// Pass F: scatter the further members into their group's CSR range.
__global__ void __launch_bounds__(256) build_dup_fill_kernel(const BuildArgs a) {
  pdl_enter();
  const uint64_t n = *a.n_dup;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 d = a.dups[i];
    a.items[slot_rec_ptr(a.tab, d.x)->z + d.w] = make_uint2(d.y, d.z);
  }
}

}  // namespace msc
