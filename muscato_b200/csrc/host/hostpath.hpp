// hostpath.hpp -- host side of the stage executables: flat parsing of the reference's text files,
// the tiled / multi-device run over the C ABI, and the multi-threaded writers of matches.txt.sz,
// results.txt and the non-match fastq (the epilogue of cmd/muscato/main.go:507-676 and
// cmd/muscato_nonmatch/main.go:57-113).  Everything is sized for BASELINE configs[2]/[3]: no
// per-record std::string, no global sort of text lines.
//
// Ordering without a text sort.  results.txt / matches.txt are bytewise-sorted lines whose first
// field is the read sequence.  reads_sorted.txt is itself bytewise sorted and duplicate free
// (`sort | muscato_uniqify`, cmd/muscato/main.go:152-221) and a tab sorts below every base letter,
// so line order on the first field IS read-id order, which is the order the library returns.
// Inside one read the remaining fields decide: target subsequence (bytes), pos and nx as DECIMAL
// STRINGS followed by a tab, then the gene column(s).  Read groups are sorted
// independently (they hold ~1 line on typical data), in parallel over ranges of reads.  If the
// read file violates the sorted-unique contract the writers fall back to a whole-line sort.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <sys/stat.h>

#include "../../../include/muscato_b200.h"
#include "szio.hpp"

namespace hostpath {

inline unsigned hw_threads() {
  unsigned n = std::thread::hardware_concurrency();
  return n ? std::min(n, 64u) : 4u;
}

// fn(part, n_parts) on n_parts threads; exceptions are re-thrown on the caller's thread.
inline void parallel_parts(unsigned n_parts, const std::function<void(unsigned, unsigned)>& fn) {
  if (n_parts <= 1) {
    fn(0, 1);
    return;
  }
  std::vector<std::thread> th;
  std::vector<std::string> errs(n_parts);
  std::atomic<bool> failed{false};
  for (unsigned t = 0; t < n_parts; t++)
    th.emplace_back([&, t] {
      try {
        fn(t, n_parts);
      } catch (const std::exception& e) {
        errs[t] = e.what();
        failed = true;
      }
    });
  for (auto& x : th) x.join();
  if (failed)
    for (auto& e : errs)
      if (!e.empty()) throw std::runtime_error(e);
}

struct Span {
  uint64_t off = 0;
  uint32_t len = 0;
};

inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\v' || c == '\f' || c == '\r' || c == '\n'; }

// Cut `text` into n pieces that end on line boundaries.
inline std::vector<size_t> line_cuts(const std::string& text, unsigned n) {
  std::vector<size_t> cut(n + 1, text.size());
  cut[0] = 0;
  for (unsigned i = 1; i < n; i++) {
    size_t p = text.size() / n * i;
    if (p < cut[i - 1]) p = cut[i - 1];
    const size_t nl = text.find('\n', p);
    cut[i] = nl == std::string::npos ? text.size() : nl + 1;
  }
  return cut;
}

// ---------------------------------------------------------------------------------------------
// reads_sorted.txt.sz: `seq \t count \t names` (cmd/muscato_uniqify/main.go:89-110); the screen and
// window_reads take bytes.Fields(line)[0] (cmd/muscato_screen/main.go:172).
// ---------------------------------------------------------------------------------------------
struct ReadSet {
  std::string text;                  // the decoded file; count / names point into it
  std::vector<Span> count, names;
  std::string ascii;                 // all sequences back to back
  std::vector<uint64_t> offs;        // n + 1
  bool sorted_unique = true;         // the file honours the `sort | uniqify` contract
  uint64_t n() const { return offs.empty() ? 0 : offs.size() - 1; }
  const char* seq(uint64_t i) const { return ascii.data() + offs[i]; }
  uint32_t len(uint64_t i) const { return (uint32_t)(offs[i + 1] - offs[i]); }
};

inline void parse_reads_sorted(ReadSet& rs, unsigned threads) {
  const std::string& text = rs.text;
  const std::vector<size_t> cut = line_cuts(text, threads);
  struct Part {
    std::vector<Span> seq, count, names;
    uint64_t bytes = 0;
  };
  std::vector<Part> parts(threads);
  parallel_parts(threads, [&](unsigned t, unsigned) {
    Part& P = parts[t];
    size_t i = cut[t];
    const size_t end = cut[t + 1];
    while (i < end) {
      size_t j = text.find('\n', i);
      if (j == std::string::npos || j > end) j = end;
      size_t n = j - i;
      if (n && text[i + n - 1] == '\r') n--;
      if (n) {
        const char* p = text.data() + i;
        size_t t1 = 0;
        while (t1 < n && p[t1] != '\t') t1++;
        size_t t2 = t1 + 1;
        while (t2 < n && p[t2] != '\t') t2++;
        size_t a = 0;
        while (a < t1 && is_ws(p[a])) a++;
        size_t b = a;
        while (b < t1 && !is_ws(p[b])) b++;
        Span s, c, nm;
        s.off = i + a;
        s.len = (uint32_t)(b - a);
        if (t1 < n) {
          c.off = i + t1 + 1;
          c.len = (uint32_t)(std::min(t2, n) - t1 - 1);
        }
        if (t2 < n) {
          nm.off = i + t2 + 1;
          nm.len = (uint32_t)(n - t2 - 1);
        }
        P.seq.push_back(s);
        P.count.push_back(c);
        P.names.push_back(nm);
        P.bytes += s.len;
      }
      i = j + 1;
    }
  });
  uint64_t n = 0, bytes = 0;
  std::vector<uint64_t> n0(threads + 1, 0), b0(threads + 1, 0);
  for (unsigned t = 0; t < threads; t++) {
    n0[t] = n;
    b0[t] = bytes;
    n += parts[t].seq.size();
    bytes += parts[t].bytes;
  }
  rs.count.resize(n);
  rs.names.resize(n);
  rs.offs.assign(n + 1, 0);
  rs.ascii.resize(bytes);
  std::vector<char> ok(threads, 1);
  parallel_parts(threads, [&](unsigned t, unsigned) {
    const Part& P = parts[t];
    uint64_t at = b0[t];
    for (size_t k = 0; k < P.seq.size(); k++) {
      const uint64_t id = n0[t] + k;
      rs.offs[id] = at;
      memcpy(&rs.ascii[at], text.data() + P.seq[k].off, P.seq[k].len);
      at += P.seq[k].len;
      rs.count[id] = P.count[k];
      rs.names[id] = P.names[k];
    }
  });
  rs.offs[n] = bytes;
  // sorted + unique on the sequence column (a proper prefix first)?
  parallel_parts(threads, [&](unsigned t, unsigned nt) {
    const uint64_t lo = n * t / nt, hi = n * (t + 1) / nt;
    for (uint64_t i = std::max<uint64_t>(lo, 1); i < hi; i++) {
      const uint32_t la = rs.len(i - 1), lb = rs.len(i);
      const int c = memcmp(rs.seq(i - 1), rs.seq(i), std::min(la, lb));
      if (c > 0 || (c == 0 && la >= lb)) {
        ok[t] = 0;
        break;
      }
    }
  });
  for (char c : ok)
    if (!c) rs.sorted_unique = false;
}

// ---------------------------------------------------------------------------------------------
// Targets: one sequence per line, the text before the first tab (cmd/muscato_screen/main.go:448-449);
// gene id = 0-based line index (:440-452).  Either ASCII (parsed from the text file) or the packed
// cache `<GeneFileName>.2bit` (f4), from which single bases are decoded for the output columns.
// ---------------------------------------------------------------------------------------------
struct TargetSet {
  std::vector<uint64_t> offs;        // n + 1 base offsets in the concatenated stream
  std::string ascii;                 // empty when loaded from the cache
  std::vector<uint64_t> words, xplane;  // packed form (cache); xplane empty = no X anywhere
  bool packed = false;
  uint64_t n() const { return offs.empty() ? 0 : offs.size() - 1; }
  uint64_t bases() const { return offs.empty() ? 0 : offs.back(); }
  inline char base(uint64_t i) const {
    if (!packed) return ascii[i];
    const uint64_t w = i >> 5;
    const unsigned sh = (unsigned)(i & 31u) * 2u;
    if (!xplane.empty() && ((xplane[w] >> sh) & 1ull)) return 'X';
    return "ACTG"[(words[w] >> sh) & 3ull];
  }
  void copy(uint64_t from, uint32_t n, char* dst) const {
    if (!packed) {
      memcpy(dst, ascii.data() + from, n);
      return;
    }
    for (uint32_t i = 0; i < n; i++) dst[i] = base(from + i);
  }
};

inline void parse_targets(const std::string& text, TargetSet& ts, unsigned threads) {
  const std::vector<size_t> cut = line_cuts(text, threads);
  struct Part {
    std::vector<Span> seq;
    uint64_t bytes = 0;
  };
  std::vector<Part> parts(threads);
  parallel_parts(threads, [&](unsigned t, unsigned) {
    size_t i = cut[t];
    const size_t end = cut[t + 1];
    while (i < end) {
      size_t j = text.find('\n', i);
      if (j == std::string::npos || j > end) j = end;
      size_t n = j - i;
      if (n && text[i + n - 1] == '\r') n--;
      size_t k = 0;
      while (k < n && text[i + k] != '\t') k++;
      Span s;
      s.off = i;
      s.len = (uint32_t)k;
      if (k != n || n > 0xffffffffull) {
        // (a tab ends the sequence column; lines are < 4 GiB by the reference's own 1 MiB scanner limit, Q12)
      }
      parts[t].seq.push_back(s);
      parts[t].bytes += k;
      i = j + 1;
    }
  });
  uint64_t n = 0, bytes = 0;
  std::vector<uint64_t> n0(threads + 1, 0), b0(threads + 1, 0);
  for (unsigned t = 0; t < threads; t++) {
    n0[t] = n;
    b0[t] = bytes;
    n += parts[t].seq.size();
    bytes += parts[t].bytes;
  }
  ts.offs.assign(n + 1, 0);
  ts.ascii.resize(bytes);
  parallel_parts(threads, [&](unsigned t, unsigned) {
    uint64_t at = b0[t];
    for (size_t k = 0; k < parts[t].seq.size(); k++) {
      ts.offs[n0[t] + k] = at;
      memcpy(&ts.ascii[at], text.data() + parts[t].seq[k].off, parts[t].seq[k].len);
      at += parts[t].seq[k].len;
    }
  });
  ts.offs[n] = bytes;
  ts.packed = false;
}

// Gene id file: `%011d \t name \t len` per target (cmd/muscato_prep_targets/main.go:296-316).
struct GeneIds {
  std::string text;
  std::vector<Span> name, len;
};

inline void parse_gene_ids(GeneIds& g) {
  const std::string& text = g.text;
  size_t i = 0;
  while (i < text.size()) {
    size_t j = text.find('\n', i);
    if (j == std::string::npos) j = text.size();
    size_t n = j - i;
    if (n && text[i + n - 1] == '\r') n--;
    if (n) {
      const char* p = text.data() + i;
      size_t t1 = 0;
      while (t1 < n && p[t1] != '\t') t1++;
      size_t t2 = t1 + 1;
      while (t2 < n && p[t2] != '\t') t2++;
      Span nm, ln;
      if (t1 < n) {
        nm.off = i + t1 + 1;
        nm.len = (uint32_t)(std::min(t2, n) - t1 - 1);
      }
      if (t2 < n) {
        ln.off = i + t2 + 1;
        ln.len = (uint32_t)(n - t2 - 1);
      }
      g.name.push_back(nm);
      g.len.push_back(ln);
    }
    i = j + 1;
  }
}

// ---------------------------------------------------------------------------------------------
// Packed target cache `<GeneFileName>.2bit`.  Header + offsets + words
// (+ X plane when any target contains X).  Valid while the source's size and mtime are unchanged.
// ---------------------------------------------------------------------------------------------
struct CacheHeader {
  char magic[8];        // "MSC2BIT1"
  uint64_t n_targets, n_bases, n_words, has_x, src_size, src_mtime_ns;
};

inline bool stat_file(const std::string& path, uint64_t& size, uint64_t& mtime_ns) {
  struct stat st;
  if (stat(path.c_str(), &st) != 0) return false;
  size = (uint64_t)st.st_size;
  mtime_ns = (uint64_t)st.st_mtim.tv_sec * 1000000000ull + (uint64_t)st.st_mtim.tv_nsec;
  return true;
}

inline bool load_target_cache(const std::string& src, const std::string& cache, TargetSet& ts) {
  uint64_t ssz = 0, smt = 0;
  if (!stat_file(src, ssz, smt)) return false;
  FILE* f = fopen(cache.c_str(), "rb");
  if (!f) return false;
  CacheHeader h;
  bool ok = fread(&h, sizeof h, 1, f) == 1 && !memcmp(h.magic, "MSC2BIT1", 8) && h.src_size == ssz && h.src_mtime_ns == smt &&
            h.n_words == (h.n_bases + 31) / 32;
  if (ok) {
    ts.offs.resize(h.n_targets + 1);
    ts.words.resize(h.n_words);
    ok = fread(ts.offs.data(), 8, ts.offs.size(), f) == ts.offs.size() &&
         (h.n_words == 0 || fread(ts.words.data(), 8, h.n_words, f) == h.n_words);
    if (ok && h.has_x) {
      ts.xplane.resize(h.n_words);
      ok = h.n_words == 0 || fread(ts.xplane.data(), 8, h.n_words, f) == h.n_words;
    } else {
      ts.xplane.clear();
    }
    ok = ok && ts.offs.back() == h.n_bases;
  }
  fclose(f);
  if (ok) {
    ts.packed = true;
    ts.ascii.clear();
  }
  return ok;
}

inline void write_target_cache(const std::string& src, const std::string& cache, const std::vector<uint64_t>& offs,
                               const std::vector<uint64_t>& words, const std::vector<uint64_t>& xplane, bool has_x) {
  CacheHeader h;
  memcpy(h.magic, "MSC2BIT1", 8);
  h.n_targets = offs.size() - 1;
  h.n_bases = offs.back();
  h.n_words = words.size();
  h.has_x = has_x ? 1 : 0;
  if (!stat_file(src, h.src_size, h.src_mtime_ns)) return;
  const std::string tmp = cache + ".tmp";
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return;  // the cache is an optimisation: a read-only directory is not an error
  bool ok = fwrite(&h, sizeof h, 1, f) == 1 && fwrite(offs.data(), 8, offs.size(), f) == offs.size() &&
            (words.empty() || fwrite(words.data(), 8, words.size(), f) == words.size());
  if (ok && has_x) ok = words.empty() || fwrite(xplane.data(), 8, words.size(), f) == words.size();
  ok = fclose(f) == 0 && ok;
  if (ok) rename(tmp.c_str(), cache.c_str());
  else remove(tmp.c_str());
}

// ---------------------------------------------------------------------------------------------
// Tiling limits of one context call (include/muscato_b200.h).
// ---------------------------------------------------------------------------------------------
constexpr uint64_t kMaxItems = 1ull << 30;
constexpr uint64_t kMaxBases = (1ull << 32) - 8192;

inline std::vector<std::pair<uint64_t, uint64_t>> split_even(uint64_t n, uint64_t parts) {
  std::vector<std::pair<uint64_t, uint64_t>> out;
  parts = std::max<uint64_t>(1, parts);
  for (uint64_t i = 0; i < parts; i++) out.emplace_back(n * i / parts, n * (i + 1) / parts);
  return out;
}

inline std::vector<std::pair<uint64_t, uint64_t>> split_target_ranges(const std::vector<uint64_t>& offs, uint64_t max_bases) {
  std::vector<std::pair<uint64_t, uint64_t>> out;
  const uint64_t G = offs.size() - 1;
  uint64_t lo = 0;
  while (lo < G) {
    const uint64_t limit = offs[lo] + max_bases;
    uint64_t hi = (uint64_t)(std::upper_bound(offs.begin(), offs.end(), limit) - offs.begin()) - 1;
    if (hi <= lo) throw std::runtime_error("a single target exceeds the per-call base limit");
    hi = std::min(hi, G);
    out.emplace_back(lo, hi);
    lo = hi;
  }
  if (out.empty()) out.emplace_back(0, 0);
  return out;
}

// Decimal string of v followed by a tab, compared bytewise ("10\t" < "9\t", "1\t" < "10\t").
inline int cmp_decimal_tab(uint32_t a, uint32_t b) {
  if (a == b) return 0;
  char sa[16], sb[16];
  const int na = snprintf(sa, sizeof sa, "%u\t", a), nb = snprintf(sb, sizeof sb, "%u\t", b);
  const int c = memcmp(sa, sb, (size_t)std::min(na, nb));
  if (c) return c;
  return na < nb ? -1 : 1;
}

inline void append_u32(std::string& s, uint32_t v) {
  char b[16];
  const int n = snprintf(b, sizeof b, "%u", v);
  s.append(b, (size_t)n);
}

// Frame a text buffer as Snappy "stored" chunks (no stream identifier): the pieces of several
// threads are concatenated behind one identifier (cmd/muscato/main.go:471-475 pipes through `sztool -c`).
inline void frame_chunks(const std::string& data, std::string& out) {
  const size_t kBlock = 65536;
  out.reserve(out.size() + data.size() + (data.size() / kBlock + 1) * 8);
  for (size_t i = 0; i < data.size(); i += kBlock) {
    const size_t n = std::min(kBlock, data.size() - i);
    const uint8_t* p = reinterpret_cast<const uint8_t*>(data.data()) + i;
    const uint32_t crc = szio::masked_crc(p, n);
    const uint32_t len = (uint32_t)n + 4;
    const char hdr[8] = {0x01, (char)len, (char)(len >> 8), (char)(len >> 16), (char)crc, (char)(crc >> 8), (char)(crc >> 16), (char)(crc >> 24)};
    out.append(hdr, 8);
    out.append(reinterpret_cast<const char*>(p), n);
  }
}

inline void write_pieces(const std::string& path, std::vector<std::string>& pieces, bool framed, unsigned threads) {
  if (framed) {
    std::vector<std::string> fr(pieces.size());
    std::atomic<size_t> next{0};
    parallel_parts(std::min<unsigned>(threads, (unsigned)std::max<size_t>(1, pieces.size())), [&](unsigned, unsigned) {
      for (size_t i = next++; i < pieces.size(); i = next++) {
        frame_chunks(pieces[i], fr[i]);
        std::string().swap(pieces[i]);
      }
    });
    pieces.swap(fr);
  }
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) throw std::runtime_error("cannot create " + path);
  if (framed) fwrite("\xff\x06\x00\x00sNaPpY", 1, 10, f);
  for (auto& p : pieces) {
    if (!p.empty() && fwrite(p.data(), 1, p.size(), f) != p.size()) {
      fclose(f);
      throw std::runtime_error("write failed: " + path);
    }
  }
  if (fclose(f) != 0) throw std::runtime_error("write failed: " + path);
}

// ---------------------------------------------------------------------------------------------
// The writers.  m[] is ordered by (read_id, gene_id, pos) with global ids.
// ---------------------------------------------------------------------------------------------
struct Outputs {
  const ReadSet* reads = nullptr;
  const TargetSet* targets = nullptr;
  const GeneIds* genes = nullptr;  // may be null (matches.txt only)
  unsigned threads = 4;
};

// Cut [0, n) into pieces on read-group boundaries.
inline std::vector<uint64_t> group_cuts(const msc_match* m, uint64_t n, unsigned pieces) {
  std::vector<uint64_t> cut(pieces + 1, n);
  cut[0] = 0;
  for (unsigned i = 1; i < pieces; i++) {
    uint64_t p = std::max(cut[i - 1], n * i / pieces);
    while (p < n && p > 0 && m[p].read_id == m[p - 1].read_id) p++;
    cut[i] = p;
  }
  return cut;
}

// One output line of matches.txt (with_names = false: `read \t tsub \t pos \t nx \t %011d gene`,
// cmd/muscato_confirm/main.go:221-230) or results.txt (`read \t tsub \t pos \t nx \t gene name \t gene len \t
// count \t names`, README.md:77-94 after cmd/muscato/main.go:574-676).
inline void format_line(const Outputs& o, const msc_match& x, bool with_names, std::string& out, std::string& scratch) {
  const ReadSet& R = *o.reads;
  const TargetSet& T = *o.targets;
  const uint32_t L = R.len(x.read_id);
  out.append(R.seq(x.read_id), L);
  out += '\t';
  scratch.resize(L);
  T.copy(T.offs[x.gene_id] + x.pos, L, &scratch[0]);
  out.append(scratch);
  out += '\t';
  append_u32(out, x.pos);
  out += '\t';
  append_u32(out, x.nx);
  out += '\t';
  if (!with_names) {
    char b[16];
    const int n = snprintf(b, sizeof b, "%011u", x.gene_id);
    out.append(b, (size_t)n);
  } else {
    const GeneIds& G = *o.genes;
    out.append(G.text, G.name[x.gene_id].off, G.name[x.gene_id].len);
    out += '\t';
    out.append(G.text, G.len[x.gene_id].off, G.len[x.gene_id].len);
    out += '\t';
    out.append(R.text, R.count[x.read_id].off, R.count[x.read_id].len);
    out += '\t';
    out.append(R.text, R.names[x.read_id].off, R.names[x.read_id].len);
  }
  out += '\n';
}

// Bytewise order of two lines of the SAME read (see the file comment).
inline bool line_less_same_read(const Outputs& o, const msc_match& a, const msc_match& b, bool with_names) {
  const TargetSet& T = *o.targets;
  const uint32_t L = o.reads->len(a.read_id);
  const uint64_t pa = T.offs[a.gene_id] + a.pos, pb = T.offs[b.gene_id] + b.pos;
  if (pa != pb) {
    for (uint32_t i = 0; i < L; i++) {
      const char ca = T.base(pa + i), cb = T.base(pb + i);
      if (ca != cb) return ca < cb;
    }
  }
  int c = cmp_decimal_tab(a.pos, b.pos);
  if (c) return c < 0;
  c = cmp_decimal_tab(a.nx, b.nx);
  if (c) return c < 0;
  if (!with_names) return a.gene_id < b.gene_id;  // %011d
  const GeneIds& G = *o.genes;
  // "name \t len \t ..." : compare name + '\t', then len + '\t'
  auto cmp_span_tab = [&](const Span& x, const Span& y) {
    const uint32_t n = std::min(x.len, y.len);
    const int r = memcmp(G.text.data() + x.off, G.text.data() + y.off, n);
    if (r) return r;
    if (x.len == y.len) return 0;
    // the shorter one continues with '\t'
    if (x.len < y.len) return (int)'\t' - (int)(unsigned char)G.text[y.off + n];
    return (int)(unsigned char)G.text[x.off + n] - (int)'\t';
  };
  c = cmp_span_tab(G.name[a.gene_id], G.name[b.gene_id]);
  if (c) return c < 0;
  c = cmp_span_tab(G.len[a.gene_id], G.len[b.gene_id]);
  if (c) return c < 0;
  return a.gene_id < b.gene_id;
}

inline void write_match_lines(const Outputs& o, const msc_match* m, uint64_t n, bool with_names, const std::string& path, bool framed) {
  const unsigned pieces_n = std::max(1u, std::min<unsigned>(o.threads * 4, (unsigned)std::max<uint64_t>(1, n / 4096)));
  std::vector<std::string> pieces(pieces_n);
  if (o.reads->sorted_unique) {
    const std::vector<uint64_t> cut = group_cuts(m, n, pieces_n);
    std::atomic<unsigned> next{0};
    parallel_parts(std::min(o.threads, pieces_n), [&](unsigned, unsigned) {
      std::vector<msc_match> grp;
      std::string scratch;
      for (unsigned pc = next++; pc < pieces_n; pc = next++) {
        std::string& out = pieces[pc];
        uint64_t i = cut[pc];
        const uint64_t end = cut[pc + 1];
        if (end > i) out.reserve((size_t)((end - i) * (2ull * o.reads->len(m[i].read_id) + 48)));
        while (i < end) {
          uint64_t j = i + 1;
          while (j < end && m[j].read_id == m[i].read_id) j++;
          if (j - i == 1) {
            format_line(o, m[i], with_names, out, scratch);
          } else {
            grp.assign(m + i, m + j);
            std::sort(grp.begin(), grp.end(), [&](const msc_match& a, const msc_match& b) { return line_less_same_read(o, a, b, with_names); });
            for (const msc_match& x : grp) format_line(o, x, with_names, out, scratch);
          }
          i = j;
        }
      }
    });
  } else {
    // contract violated (reads not sorted / not unique): whole-line sort, as `sort` would do
    std::vector<std::string> lines(n);
    std::string scratch;
    for (uint64_t i = 0; i < n; i++) format_line(o, m[i], with_names, lines[i], scratch);
    if (with_names) {
      // `sort -k1` runs BEFORE the read columns are joined (cmd/muscato/main.go:640-676): order on the
      // first six fields only -- the count/names columns are identical for equal reads anyway
    }
    std::sort(lines.begin(), lines.end());
    pieces.assign(1, std::string());
    for (auto& l : lines) pieces[0] += l;
  }
  write_pieces(path, pieces, framed, o.threads);
}

// Non-match fastq (cmd/muscato_nonmatch/main.go:95-113): for every read id in `ids` (ascending =
// reads_sorted order): first whitespace token of names + '#' + count, sequence, '+', '!' x L.
inline void write_nonmatch(const Outputs& o, const uint32_t* ids, uint64_t n, const std::string& path) {
  const ReadSet& R = *o.reads;
  const unsigned pieces_n = std::max(1u, std::min<unsigned>(o.threads * 4, (unsigned)std::max<uint64_t>(1, n / 4096)));
  std::vector<std::string> pieces(pieces_n);
  std::atomic<unsigned> next{0};
  parallel_parts(std::min(o.threads, pieces_n), [&](unsigned, unsigned) {
    for (unsigned pc = next++; pc < pieces_n; pc = next++) {
      std::string& out = pieces[pc];
      const uint64_t lo = n * pc / pieces_n, hi = n * (pc + 1) / pieces_n;
      for (uint64_t k = lo; k < hi; k++) {
        const uint32_t i = ids[k];
        const char* nm = R.text.data() + R.names[i].off;
        const uint32_t nl = R.names[i].len;
        uint32_t a = 0;
        while (a < nl && is_ws(nm[a])) a++;
        uint32_t b = a;
        while (b < nl && !is_ws(nm[b])) b++;
        out.append(nm + a, b - a);
        out += '#';
        out.append(R.text, R.count[i].off, R.count[i].len);
        out += '\n';
        out.append(R.seq(i), R.len(i));
        out += "\n+\n";
        out.append(R.len(i), '!');
        out += '\n';
      }
    }
  });
  write_pieces(path, pieces, false, o.threads);
}

}  // namespace hostpath
