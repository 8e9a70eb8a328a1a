// gendat.cc -- block-structured synthetic inputs in the shape of muscato_gendat
// (cmd/muscato_gendat/main.go:39-136: iid uniform A/T/G/C reads and genes, reads copied into genes)
// for the large BASELINE configurations (S2 / S3 / S4 of SURVEY.md 8d).  Host code, no CUDA: bench
// and test tooling, not part of libmuscato_b200.so.
//
// The workload is a sequence of independent BLOCKS.  Block b holds `genes_per_block` forward genes
// (with rev: each followed by its reverse complement, cmd/muscato_prep_targets/main.go:115-134) and
// `reads_per_block` reads, the first `planted` of which are sampled from uniform random positions
// of the block's own targets with per-base substitution probability sub256/256 (always to a
// different base).  A block depends on (seed, b) only, so any prefix of the blocks -- a CPU
// sample, a rank's shard -- is the same workload at a smaller size, with the same density of true
// alignments per read, and every process regenerates identical bytes without communication.
#include <cstdint>
#include <cstring>

namespace {

struct SplitMix {
  uint64_t s;
  explicit SplitMix(uint64_t seed) : s(seed) {}
  inline uint64_t next() {
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  inline uint64_t below(uint64_t n) { return (uint64_t)(((unsigned __int128)next() * n) >> 64); }
};

const char kBases[4] = {'A', 'T', 'G', 'C'};  // gendat order (cmd/muscato_gendat/main.go:83)

inline uint64_t stream_seed(uint64_t seed, uint64_t block, uint64_t what) {
  SplitMix m(seed * 0xD6E8FEB86659FD93ull + block * 0x9E3779B97F4A7C15ull + what);
  m.next();
  return m.next();
}

void fill_random(SplitMix& g, uint8_t* out, uint64_t n) {
  uint64_t i = 0;
  while (i + 32 <= n) {
    uint64_t z = g.next();
    for (int k = 0; k < 32; k++) out[i + k] = (uint8_t)kBases[(z >> (2 * k)) & 3];
    i += 32;
  }
  if (i < n) {
    uint64_t z = g.next();
    for (; i < n; i++, z >>= 2) out[i] = (uint8_t)kBases[z & 3];
  }
}

inline uint8_t comp(uint8_t c) {
  switch (c) {
    case 'A': return 'T';
    case 'T': return 'A';
    case 'G': return 'C';
    case 'C': return 'G';
    default: return 'X';
  }
}

inline int base_index(uint8_t c) { return c == 'A' ? 0 : c == 'T' ? 1 : c == 'G' ? 2 : 3; }

}  // namespace

extern "C" {

// Targets of block `block`: genes_per_block * (rev ? 2 : 1) sequences of gene_len bases, written
// back to back into out (2i = forward, 2i+1 = reverse complement when rev).  period > 0 makes the
// block low-complexity (S4): every gene is a tandem repeat of a random unit of pmin..pmax bases (period = pmax | pmin << 8) with
// sub256/256 of its bases substituted.
void msc_gen_targets(uint64_t seed, uint64_t block, uint32_t genes_per_block, uint32_t gene_len, int rev, uint32_t period,
                     uint32_t sub256, uint8_t* out) {
  SplitMix g(stream_seed(seed, block, 1));
  const uint64_t stride = (uint64_t)gene_len * (rev ? 2 : 1);
  for (uint32_t i = 0; i < genes_per_block; i++) {
    uint8_t* fw = out + (uint64_t)i * stride;
    if (period == 0) {
      fill_random(g, fw, gene_len);
    } else {
      uint8_t unit[64];
      // period = longest unit | shortest unit << 8 (0 = 1)
      const uint32_t pmax = (period & 255u) < 64 ? (period & 255u) : 64, pmin = (period >> 8) ? (period >> 8) : 1;
      const uint32_t p = pmin >= pmax ? pmax : pmin + (uint32_t)g.below(pmax - pmin + 1);
      fill_random(g, unit, p);
      for (uint32_t j = 0; j < gene_len; j++) fw[j] = unit[j % p];
      if (sub256) {
        for (uint32_t j = 0; j < gene_len; j++) {
          const uint64_t z = g.next();
          if ((z & 255u) < sub256) fw[j] = (uint8_t)kBases[(base_index(fw[j]) + 1 + ((z >> 8) % 3)) & 3];
        }
      }
    }
    if (rev) {
      uint8_t* rc = fw + gene_len;
      for (uint32_t j = 0; j < gene_len; j++) rc[j] = comp(fw[gene_len - 1 - j]);
    }
  }
}

// Reads of block `block` (reads_per_block x read_len, back to back).  targets = the block's own
// target buffer as msc_gen_targets wrote it (n_targets sequences of gene_len bases).  The first
// `planted` reads are copies of target[g][p : p + read_len] with substitutions; plant_gene /
// plant_pos (may be NULL) receive g (block-local target index) and p for them.
void msc_gen_reads(uint64_t seed, uint64_t block, uint32_t reads_per_block, uint32_t read_len, uint32_t planted,
                   uint32_t sub256, const uint8_t* targets, uint32_t n_targets, uint32_t gene_len, uint8_t* out,
                   int32_t* plant_gene, int32_t* plant_pos) {
  SplitMix g(stream_seed(seed, block, 2));
  if (planted > reads_per_block) planted = reads_per_block;
  if (gene_len < read_len || n_targets == 0) planted = 0;
  for (uint32_t i = 0; i < planted; i++) {
    const uint32_t t = (uint32_t)g.below(n_targets);
    const uint32_t p = (uint32_t)g.below((uint64_t)gene_len - read_len + 1);
    uint8_t* r = out + (uint64_t)i * read_len;
    memcpy(r, targets + (uint64_t)t * gene_len + p, read_len);
    if (sub256) {
      uint32_t j = 0;
      while (j < read_len) {  // 8 bases per draw: one byte decides, the next two bits pick the new base
        uint64_t z = g.next();
        uint64_t z2 = 0;
        bool have2 = false;
        for (int k = 0; k < 8 && j < read_len; k++, j++, z >>= 8) {
          if ((z & 255u) < sub256) {
            if (!have2) { z2 = g.next(); have2 = true; }
            r[j] = (uint8_t)kBases[(base_index(r[j]) + 1 + (z2 % 3)) & 3];
            z2 /= 3;
          }
        }
      }
    }
    if (plant_gene) plant_gene[i] = (int32_t)t;
    if (plant_pos) plant_pos[i] = (int32_t)p;
  }
  fill_random(g, out + (uint64_t)planted * read_len, (uint64_t)(reads_per_block - planted) * read_len);
}

}  // extern "C"
