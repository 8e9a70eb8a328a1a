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
// Data movement: target tiles (256 words = 8192 bases + halo) are staged into shared memory
// by the TMA engine (1-D cp.async.bulk + mbarrier, double buffered); a persistent grid walks
// the tiles.  The Bloom/table probes are scattered 8-byte reads served from L2 when the
// filter fits (it is sized 32-64 bits per key) and from HBM sectors otherwise.
#pragma once
#include "common.cuh"

namespace msc {

constexpr int kScanBlock = 256;
constexpr int kTileWords = 256;                       // one 32-base word per thread per tile
constexpr int kTileCopyWords = kTileWords + 2;        // halo word + 1 (byte count multiple of 16)
constexpr int kTileSmemWords = kTileWords + 8;

struct ScanArgs {
  const uint64_t* tg_words;
  const uint64_t* tg_x;
  const uint32_t* xsum;
  uint64_t n_bases;
  uint64_t n_tiles;
  const uint2* bloom;
  int lg_bloom;
  const uint64_t* tab_fp;
  int lg_slots;
  uint2* cand;
  unsigned long long cand_cap;
  unsigned long long* n_cand;
  unsigned long long* n_bloom_pass;
  int W;
};

__device__ __forceinline__ void scan_probe_emit(const ScanArgs& a, uint64_t fp, uint32_t gpos) {
  const int64_t s = table_find(a.tab_fp, a.lg_slots, fp);
  if (s >= 0) {
    const unsigned long long at = warp_agg_inc(a.n_cand);
    if (at < a.cand_cap) a.cand[at] = make_uint2((uint32_t)s, gpos);
  }
}

__global__ void __launch_bounds__(kScanBlock) scan_targets_kernel(const ScanArgs a) {
  __shared__ alignas(128) uint64_t tile[2][kTileSmemWords];
  __shared__ alignas(8) uint64_t bar[2];
  const int tid = threadIdx.x;
  constexpr uint32_t kBytes = kTileCopyWords * sizeof(uint64_t);
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  const uint64_t kmask = low_bases_mask(a.W);
  uint64_t t = blockIdx.x;
  if (tid == 0 && t < a.n_tiles) {
    mbar_arrive_expect_tx(&bar[0], kBytes);
    bulk_copy_g2s(tile[0], a.tg_words + t * kTileWords, kBytes, &bar[0]);
  }
  uint32_t phases = 0;
  int buf = 0;
  uint32_t n_pass = 0;
  for (; t < a.n_tiles; t += gridDim.x) {
    const uint64_t tn = t + gridDim.x;
    if (tid == 0 && tn < a.n_tiles) {
      mbar_arrive_expect_tx(&bar[buf ^ 1], kBytes);
      bulk_copy_g2s(tile[buf ^ 1], a.tg_words + tn * kTileWords, kBytes, &bar[buf ^ 1]);
    }
    mbar_wait(&bar[buf], (phases >> buf) & 1u);
    phases ^= 1u << buf;

    const uint64_t lo = tile[buf][tid];
    const uint64_t hi = tile[buf][tid + 1];
    const uint64_t w = t * kTileWords + (uint64_t)tid;
    const uint64_t gbase = w * 32ull;
    if (gbase < a.n_bases) {
      const int npos = (int)min((uint64_t)32, a.n_bases - gbase);
      // X summary bits of word w and w+1 (xsum is padded by one uint32).
      const uint32_t xs0 = __ldg(a.xsum + (w >> 5));
      const uint32_t xs1 = __ldg(a.xsum + ((w + 1) >> 5));
      const bool anyx = ((xs0 >> (unsigned)(w & 31u)) | (xs1 >> (unsigned)((w + 1) & 31u))) & 1u;
      if (!anyx) {
#pragma unroll
        for (int jb = 0; jb < 32; jb += 8) {
          uint64_t fp[8];
          uint2 bw[8];
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const int j = jb + i;
            const uint64_t key = (j == 0 ? lo : ((lo >> (2 * j)) | (hi << (64 - 2 * j)))) & kmask;
            fp[i] = key_fp(key, 0ull);
            bw[i] = __ldg(a.bloom + bloom_index(fp[i], a.lg_bloom));
          }
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const uint32_t mlo = bloom_mask_lo(fp[i]), mhi = bloom_mask_hi(fp[i]);
            if (((bw[i].x & mlo) == mlo) & ((bw[i].y & mhi) == mhi) & (jb + i < npos)) {
              n_pass++;
              scan_probe_emit(a, fp[i], (uint32_t)(gbase + jb + i));
            }
          }
        }
      } else {
        // Slow path: the word (or its successor) contains X; fold the X mask into the key.
        const uint64_t xlo = __ldg(a.tg_x + w);
        const uint64_t xhi = __ldg(a.tg_x + w + 1);
        for (int j = 0; j < npos; j++) {
          const uint64_t key = (j == 0 ? lo : ((lo >> (2 * j)) | (hi << (64 - 2 * j)))) & kmask;
          const uint64_t xm = (j == 0 ? xlo : ((xlo >> (2 * j)) | (xhi << (64 - 2 * j)))) & kmask;
          const uint64_t fp = key_fp(key, xm);
          const uint2 bw = __ldg(a.bloom + bloom_index(fp, a.lg_bloom));
          const uint32_t mlo = bloom_mask_lo(fp), mhi = bloom_mask_hi(fp);
          if (((bw.x & mlo) == mlo) & ((bw.y & mhi) == mhi)) {
            n_pass++;
            scan_probe_emit(a, fp, (uint32_t)(gbase + j));
          }
        }
      }
    }
    __syncthreads();  // all reads of tile[buf] are done before it is refilled
    buf ^= 1;
  }
  n_pass = __reduce_add_sync(0xffffffffu, n_pass);
  if ((tid & 31) == 0 && n_pass) atomicAdd(a.n_bloom_pass, (unsigned long long)n_pass);
}

}  // namespace msc
