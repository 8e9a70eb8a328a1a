// muscato_b200_hotpath -- stage-compatible executables for the hot path.
//
//   muscato_b200_hotpath <LogDir/config.json> [--device N | --devices a,b,..] [--from-fastq] [--no-epilogue]
//                        [--no-target-cache] [--threads N] [--max-items N] [--max-bases N] [--report FILE]
//   muscato_b200_hotpath <LogDir/config.json> --format-matches FILE   (formats an existing match set -- binary msc_match
//                        records with global ids, e.g. gathered from several ranks -- into the three output files; no GPU)
//
// Drop-in for steps 5-12 of the reference driver (cmd/muscato/main.go:1029-1051: screen, sortBloom,
// confirm, combineWindows, sortByGeneId, joinGeneNames, joinReadNames, writeNonMatch) with the
// reference's own contracts: the same config.json (utils/config.go), TempDir/reads_sorted.txt.sz,
// Config.GeneFileName, Config.GeneIdFileName in; TempDir/matches.txt.sz, Config.ResultsFileName and
// the non-match fastq out.  All matching is done by libmuscato_b200.so on the GPU through the C ABI
// (include/muscato_b200.h); this file only parses and formats text (hostpath.hpp).  Inputs of any
// size: the reads are cut into batches of at most 2^30 read-window items and the targets into
// ranges of fewer than 2^32 bases (one context call each); several devices take contiguous parts of
// the reads (every rule of the path is per read: no exchange).  Exit status != 0 on any failure,
// like the reference stages (log.Fatal / panic), upon which the driver panics (cmd/muscato/main.go:313-315).
//
// The same binary answers to the per-stage names of the unmodified driver (argv[0]):
//   muscato_screen <config.json> [tmpdir]      cmd/muscato/main.go:310    runs the whole hot path on the GPU, keeps the
//                                              result as TempDir/msc_b200_matches.txt.sz and leaves EMPTY bmatch_<k>.txt.sz
//                                              files for the driver's sortBloom step (the candidates never leave the device)
//   muscato_confirm <config.json> <k>          cmd/muscato/main.go:402    window 0 hands the result over as rmatch_0.txt.sz,
//                                              the other windows write empty files; up to MaxConfirmProcs of them run
//                                              concurrently (:391-420) -- they serialise on TempDir/msc_b200.lock and only
//                                              the first one runs the GPU path if muscato_screen has not
//   muscato_combine_windows <config.json>      cmd/muscato/main.go:442-469   stdin -> stdout, the MMTol rule of
//                                              cmd/muscato_combine_windows/main.go:36-60 (the identity on our result)
#include <fcntl.h>
#include <sys/file.h>
#include <unistd.h>

#include <chrono>
#include <mutex>

#include "hostpath.hpp"

using namespace hostpath;

namespace {

struct Cfg {
  std::string ReadFileName, GeneFileName, GeneIdFileName, ResultsFileName, TempDir, LogDir, MatchMode;
  std::vector<int> Windows;
  int WindowWidth = 0, MinDinuc = 0, MinReadLength = 0, MaxReadLength = 0, MMTol = 0;
  long long MaxMatches = 0;
  double PMatch = 0;
};

// Flat JSON object reader (what utils.ReadConfig decodes, utils/config.go:103-117).
struct Json {
  const std::string& s;
  size_t i = 0;
  explicit Json(const std::string& t) : s(t) {}
  void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) i++; }
  bool lit(char c) { ws(); if (i < s.size() && s[i] == c) { i++; return true; } return false; }
  std::string str() {
    ws();
    if (i >= s.size() || s[i] != '"') throw std::runtime_error("config: expected a string");
    std::string o;
    for (i++; i < s.size() && s[i] != '"'; i++) {
      if (s[i] == '\\' && i + 1 < s.size()) {
        const char c = s[++i];
        o += c == 'n' ? '\n' : c == 't' ? '\t' : c == 'r' ? '\r' : c;
      } else {
        o += s[i];
      }
    }
    i++;
    return o;
  }
  std::string tok() {
    ws();
    const size_t j = i;
    while (i < s.size() && !strchr(",]} \t\r\n", s[i])) i++;
    return s.substr(j, i - j);
  }
};

Cfg read_cfg(const std::string& path) {
  const std::string txt = szio::read_all(path);
  Cfg c;
  Json j(txt);
  if (!j.lit('{')) throw std::runtime_error("config: expected an object");
  while (!j.lit('}')) {
    const std::string k = j.str();
    if (!j.lit(':')) throw std::runtime_error("config: expected ':'");
    j.ws();
    if (txt[j.i] == '"') {
      const std::string v = j.str();
      if (k == "ReadFileName") c.ReadFileName = v;
      else if (k == "GeneFileName") c.GeneFileName = v;
      else if (k == "GeneIdFileName") c.GeneIdFileName = v;
      else if (k == "ResultsFileName") c.ResultsFileName = v;
      else if (k == "TempDir") c.TempDir = v;
      else if (k == "LogDir") c.LogDir = v;
      else if (k == "MatchMode") c.MatchMode = v;
    } else if (txt[j.i] == '[') {
      j.lit('[');
      std::vector<int> a;
      while (!j.lit(']')) { a.push_back(atoi(j.tok().c_str())); j.lit(','); }
      if (k == "Windows") c.Windows = a;
    } else {
      const std::string v = j.tok();
      if (k == "WindowWidth") c.WindowWidth = atoi(v.c_str());
      else if (k == "PMatch") c.PMatch = strtod(v.c_str(), nullptr);
      else if (k == "MinDinuc") c.MinDinuc = atoi(v.c_str());
      else if (k == "MinReadLength") c.MinReadLength = atoi(v.c_str());
      else if (k == "MaxReadLength") c.MaxReadLength = atoi(v.c_str());
      else if (k == "MaxMatches") c.MaxMatches = atoll(v.c_str());
      else if (k == "MMTol") c.MMTol = atoi(v.c_str());
    }
    j.lit(',');
  }
  // checkArgs (cmd/muscato/main.go:833-904)
  if (c.Windows.empty()) throw std::runtime_error("Windows not provided");
  if (c.WindowWidth == 0) throw std::runtime_error("WindowWidth not provided");
  if (c.MaxReadLength == 0) throw std::runtime_error("MaxReadLength not provided");
  if (c.PMatch == 0) c.PMatch = 1;
  if (c.MaxMatches == 0) c.MaxMatches = 1000 * 1000;
  if (c.MatchMode.empty()) c.MatchMode = "best";
  if (c.ResultsFileName.empty()) c.ResultsFileName = "results.txt";
  if (c.TempDir.empty()) throw std::runtime_error("TempDir must be set");
  return c;
}

std::string nonmatch_name(const std::string& results) {  // cmd/muscato_nonmatch/main.go:66-71
  std::string dir, base = results;
  const size_t sl = results.rfind('/');
  if (sl != std::string::npos) { dir = results.substr(0, sl + 1); base = results.substr(sl + 1); }
  std::vector<std::string> c;
  size_t i = 0;
  while (true) {
    const size_t j = base.find('.', i);
    if (j == std::string::npos) { c.push_back(base.substr(i)); break; }
    c.push_back(base.substr(i, j - i));
    i = j + 1;
  }
  const std::string d = c.back();
  c.back() = "nonmatch";
  c.push_back(d + ".fastq");
  std::string out = dir;
  for (size_t k = 0; k < c.size(); k++) { if (k) out += '.'; out += c[k]; }
  return out;
}

msc_config make_msc_config(const Cfg& cfg, int device) {
  msc_config mc;
  memset(&mc, 0, sizeof mc);
  if (cfg.Windows.size() > MSC_MAX_WINDOWS) throw std::runtime_error("more than 32 windows");
  mc.n_windows = (int32_t)cfg.Windows.size();
  for (size_t k = 0; k < cfg.Windows.size(); k++) mc.windows[k] = cfg.Windows[k];
  mc.window_width = cfg.WindowWidth;
  mc.max_read_length = cfg.MaxReadLength;
  mc.pmatch = cfg.PMatch;
  mc.min_dinuc = cfg.MinDinuc;
  mc.mmtol = cfg.MMTol;
  mc.max_matches = cfg.MaxMatches;
  if (cfg.MatchMode != "first" && cfg.MatchMode != "best") throw std::runtime_error("MatchMode must be 'first' or 'best'");
  mc.match_mode = cfg.MatchMode == "first" ? MSC_MATCH_FIRST : MSC_MATCH_BEST;
  mc.device = device;
  return mc;
}

struct Options {
  std::vector<int> devices{0};
  bool from_fastq = false, epilogue = true, target_cache = true;
  unsigned threads = hw_threads();
  uint64_t max_items = kMaxItems, max_bases = kMaxBases;
  std::string report;
  std::string matches_file;  // --format-matches: write the output files for an existing match set (no GPU work)
};

struct RunTotals {
  uint64_t n_reads = 0, target_bases = 0, candidates = 0, pairs = 0, matches = 0, tiles = 0, h2d = 0, d2h = 0;
  double ms_scan = 0, gpu_s = 0;
  std::mutex mu;
};

struct CtxGuard {
  msc_ctx* ctx = nullptr;
  ~CtxGuard() { if (ctx) msc_destroy(ctx); }
};

void check(msc_ctx* ctx, int rc, const char* what) {
  if (rc != MSC_OK) throw std::runtime_error(std::string(what) + ": " + msc_last_error(ctx));
}

// prepReads = muscato_prep_reads | sort | muscato_uniqify (cmd/muscato/main.go:152-221) with the
// sort and the collapse of equal sequences done on the GPU (msc_prep_reads).  The host parses the
// fastq records (utils/fastq.go:35-61), hands the raw sequences over, and joins counts and names
// from the grouping the library returns: inside a group the `seq\tname` lines are in bytewise
// name order, names longer than 1000 bytes are cut to 995 + "..." before sorting
// (cmd/muscato_prep_reads/main.go:76-79), the name column is the text before its first tab and
// the joined names are cut to 996 + "..." (cmd/muscato_uniqify/main.go:89-111).  Returns the text
// of reads_sorted.txt.
std::string prep_reads_device(msc_ctx* ctx, const std::string& fastq, int min_len, int max_len) {
  std::vector<std::pair<size_t, size_t>> ln;
  {
    size_t i = 0;
    while (i < fastq.size()) {
      size_t j = fastq.find('\n', i);
      if (j == std::string::npos) j = fastq.size();
      size_t n = j - i;
      if (n && fastq[i + n - 1] == '\r') n--;
      ln.emplace_back(i, n);
      i = j + 1;
    }
  }
  std::vector<std::pair<size_t, size_t>> names;
  std::string raw;
  std::vector<uint64_t> offs(1, 0);
  for (size_t i = 0; i + 4 <= ln.size(); i += 4) {
    names.push_back(ln[i]);
    raw.append(fastq, ln[i + 1].first, ln[i + 1].second);
    offs.push_back(raw.size());
  }
  const uint64_t n_raw = names.size();
  uint64_t kept = 0, uniq = 0;
  check(ctx, msc_prep_reads(ctx, reinterpret_cast<const uint8_t*>(raw.data()), offs.data(), n_raw, min_len, &kept, &uniq), "msc_prep_reads");
  std::vector<uint32_t> perm(kept + 1), gs(uniq + 1);
  check(ctx, msc_fetch_read_groups(ctx, perm.data(), gs.data()), "msc_fetch_read_groups");
  std::string txt;
  std::vector<std::string> nm;
  for (uint64_t u = 0; u < uniq; u++) {
    nm.clear();
    for (uint32_t j = gs[u]; j < gs[u + 1]; j++) {
      std::string n(fastq, names[perm[j]].first, names[perm[j]].second);
      if (n.size() > 1000) n = n.substr(0, 995) + "...";
      nm.push_back(n);
    }
    std::sort(nm.begin(), nm.end());
    std::string na;
    for (size_t i = 0; i < nm.size(); i++) {
      if (i) na += ';';
      const size_t t = nm[i].find('\t');
      na += t == std::string::npos ? nm[i] : nm[i].substr(0, t);  // bytes.Split(line, "\t")[1]
    }
    if (na.size() > 1000) na = na.substr(0, 996) + "...";
    const uint32_t rep = perm[gs[u]];
    std::string seq(raw, offs[rep], std::min<uint64_t>(offs[rep + 1] - offs[rep], (uint64_t)max_len));  // :67-69
    for (auto& ch : seq)
      if (ch != 'A' && ch != 'T' && ch != 'C' && ch != 'G') ch = 'X';  // subx :33-44
    txt += seq;
    txt += '\t';
    txt += std::to_string(nm.size());
    txt += '\t';
    txt += na;
    txt += '\n';
  }
  return txt;
}

// Matches of one read batch over several target ranges: every range was combined against ITS
// per-read minimum (a superset of the final result); keep nx <= global minimum + MMTol
// (cmd/muscato_combine_windows/main.go:36-60 is a per-read rule over all targets).
void filter_by_global_best(std::vector<msc_match>& m, int mmtol) {
  std::sort(m.begin(), m.end(), [](const msc_match& a, const msc_match& b) {
    if (a.read_id != b.read_id) return a.read_id < b.read_id;
    if (a.gene_id != b.gene_id) return a.gene_id < b.gene_id;
    return a.pos < b.pos;
  });
  size_t w = 0;
  for (size_t i = 0; i < m.size();) {
    size_t j = i;
    uint32_t best = 0xffffffffu;
    while (j < m.size() && m[j].read_id == m[i].read_id) best = std::min(best, m[j++].nx);
    for (size_t k = i; k < j; k++)
      if (m[k].nx <= best + (uint32_t)mmtol) m[w++] = m[k];
    i = j;
  }
  m.resize(w);
}

// One device: its contiguous part of the reads against all target ranges.
void run_device(const Cfg& cfg, const Options& opt, int device, const ReadSet& reads, uint64_t r_lo, uint64_t r_hi,
                const TargetSet& targets, const std::vector<std::pair<uint64_t, uint64_t>>& ranges, bool write_cache,
                std::vector<msc_match>& out, std::vector<uint32_t>& nonmatch, bool& nonmatch_valid, RunTotals& tot) {
  const auto t0 = std::chrono::steady_clock::now();
  msc_config mc = make_msc_config(cfg, device);
  char err[512] = {0};
  CtxGuard g;
  g.ctx = msc_create(&mc, err, sizeof err);
  if (!g.ctx) throw std::runtime_error(std::string("msc_create: ") + err);
  msc_ctx* ctx = g.ctx;
  if (ranges.size() > 1) check(ctx, msc_set_shards(ctx, (int32_t)ranges.size()), "msc_set_shards");
  const uint64_t nwin = cfg.Windows.size();
  const uint64_t per = std::max<uint64_t>(1, opt.max_items / nwin);
  const uint64_t n_part = r_hi - r_lo;
  const uint64_t n_batches = std::max<uint64_t>(1, (n_part + per - 1) / per);
  nonmatch_valid = ranges.size() == 1;
  std::vector<uint64_t> roffs, toffs;
  for (uint64_t b = 0; b < n_batches; b++) {
    const uint64_t b_lo = r_lo + n_part * b / n_batches, b_hi = r_lo + n_part * (b + 1) / n_batches;
    const uint64_t nb = b_hi - b_lo;
    roffs.resize(nb + 1);
    for (uint64_t i = 0; i <= nb; i++) roffs[i] = reads.offs[b_lo + i] - reads.offs[b_lo];
    check(ctx, msc_set_reads(ctx, reinterpret_cast<const uint8_t*>(reads.ascii.data()) + reads.offs[b_lo], roffs.data(), nb), "msc_set_reads");
    std::vector<msc_match> batch;
    for (size_t ri = 0; ri < ranges.size(); ri++) {
      const uint64_t g_lo = ranges[ri].first, g_hi = ranges[ri].second, ng = g_hi - g_lo;
      toffs.resize(ng + 1);
      for (uint64_t i = 0; i <= ng; i++) toffs[i] = targets.offs[g_lo + i] - targets.offs[g_lo];
      if (targets.packed) {
        // the cache holds the whole database as ONE stream: it is only used for single-range databases
        check(ctx, msc_set_targets_packed(ctx, targets.words.data(), targets.xplane.empty() ? nullptr : targets.xplane.data(), toffs.data(), ng),
              "msc_set_targets_packed");
      } else {
        check(ctx, msc_set_targets(ctx, reinterpret_cast<const uint8_t*>(targets.ascii.data()) + targets.offs[g_lo], toffs.data(), ng),
              "msc_set_targets");
        if (write_cache && b == 0 && ranges.size() == 1) {
          std::vector<uint64_t> words(msc_packed_target_words(ctx)), xp(words.size());
          int32_t has_x = 0;
          check(ctx, msc_fetch_packed_targets(ctx, words.data(), xp.data(), &has_x), "msc_fetch_packed_targets");
          write_target_cache(cfg.GeneFileName, cfg.GeneFileName + ".2bit", targets.offs, words, xp, has_x != 0);
        }
      }
      check(ctx, msc_run(ctx), "msc_run");
      if (ranges.size() > 1 && msc_shard_overflow(ctx))
        throw std::runtime_error("a key group may exceed MaxMatches across target ranges of more than 2^32 bases: not supported by "
                                 "the stage executable (use the sharded protocol of the library, include/muscato_b200.h)");
      msc_stats st;
      msc_get_stats(ctx, &st);
      const size_t at = batch.size();
      batch.resize(at + st.n_matches);
      uint64_t n = 0;
      check(ctx, msc_fetch_matches_into(ctx, batch.data() + at, st.n_matches, &n), "msc_fetch_matches_into");
      for (size_t i = at; i < batch.size(); i++) {
        batch[i].read_id += (uint32_t)b_lo;
        batch[i].gene_id += (uint32_t)g_lo;
      }
      if (nonmatch_valid) {
        uint64_t nn = 0;
        check(ctx, msc_fetch_nonmatch(ctx, nullptr, 0, &nn), "msc_fetch_nonmatch");
        const size_t na = nonmatch.size();
        nonmatch.resize(na + nn);
        if (nn) check(ctx, msc_fetch_nonmatch(ctx, nonmatch.data() + na, nn, &nn), "msc_fetch_nonmatch");
        for (size_t i = na; i < nonmatch.size(); i++) nonmatch[i] += (uint32_t)b_lo;
      }
      std::lock_guard<std::mutex> lk(tot.mu);
      tot.candidates += st.n_candidates;
      tot.pairs += st.n_pairs;
      tot.tiles++;
    }
    if (ranges.size() > 1) filter_by_global_best(batch, cfg.MMTol);
    out.insert(out.end(), batch.begin(), batch.end());
  }
  msc_stats st;
  msc_get_stats(ctx, &st);
  std::lock_guard<std::mutex> lk(tot.mu);
  tot.h2d += st.h2d_bytes;
  tot.d2h += st.d2h_bytes;
  tot.ms_scan += st.ms_scan;
  tot.gpu_s = std::max(tot.gpu_s, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
}

// The fused hot path + epilogue.  keep_copy: also store the matches as TempDir/msc_b200_matches.txt.sz (stage shims).
int run_hotpath(const Cfg& cfg, const Options& opt, bool keep_copy) {
  const auto t_start = std::chrono::steady_clock::now();
  auto since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count(); };
  // ---- targets (packed cache, or the text file) ------------------------------------------------
  TargetSet targets;
  const std::string cache = cfg.GeneFileName + ".2bit";
  bool from_cache = false, write_cache = false;
  if (opt.target_cache && load_target_cache(cfg.GeneFileName, cache, targets) && targets.bases() < opt.max_bases) {
    from_cache = true;
  } else {
    targets = TargetSet();
    parse_targets(szio::read_text(cfg.GeneFileName), targets, opt.threads);
    write_cache = opt.target_cache;
  }
  const auto ranges = split_target_ranges(targets.offs, opt.max_bases);
  const double t_targets = since(t_start);

  // ---- reads ------------------------------------------------------------------------------------
  const auto t_r0 = std::chrono::steady_clock::now();
  ReadSet reads;
  const std::string rs_path = cfg.TempDir + "/reads_sorted.txt.sz";
  if (opt.from_fastq) {
    // prepReads with the sort / uniqify on the device (one context call: up to 2^32 - 1 raw reads)
    msc_config mc = make_msc_config(cfg, opt.devices[0]);
    char err[512] = {0};
    CtxGuard g;
    g.ctx = msc_create(&mc, err, sizeof err);
    if (!g.ctx) throw std::runtime_error(std::string("msc_create: ") + err);
    reads.text = prep_reads_device(g.ctx, szio::read_all(cfg.ReadFileName), cfg.MinReadLength, cfg.MaxReadLength);
    szio::write_file(rs_path, reads.text, true);
  } else {
    reads.text = szio::read_text(rs_path);
  }
  parse_reads_sorted(reads, opt.threads);
  const double t_reads = since(t_r0);

  // ---- the GPU path: every device takes a contiguous part of the reads -------------------------------
  const auto t_g0 = std::chrono::steady_clock::now();
  const size_t nd = opt.matches_file.empty() ? opt.devices.size() : 0;
  const auto rparts = split_even(reads.n(), std::max<size_t>(nd, 1));
  std::vector<std::vector<msc_match>> dm(std::max<size_t>(nd, 1));
  std::vector<std::vector<uint32_t>> dn(std::max<size_t>(nd, 1));
  std::vector<char> nm_valid(std::max<size_t>(nd, 1), 0);
  RunTotals tot;
  if (nd) parallel_parts((unsigned)nd, [&](unsigned d, unsigned) {
    bool v = false;
    run_device(cfg, opt, opt.devices[d], reads, rparts[d].first, rparts[d].second, targets, ranges, write_cache && d == 0, dm[d], dn[d], v, tot);
    nm_valid[d] = v ? 1 : 0;
  });
  std::vector<msc_match> m;
  if (nd == 0) {
    const std::string raw = szio::read_all(opt.matches_file);
    if (raw.size() % sizeof(msc_match)) throw std::runtime_error("--format-matches: file is not a whole number of records");
    m.resize(raw.size() / sizeof(msc_match));
    if (!m.empty()) memcpy(m.data(), raw.data(), raw.size());
    for (size_t i = 0; i < m.size(); i++) {
      if (m[i].read_id >= reads.n() || m[i].gene_id >= targets.n()) throw std::runtime_error("--format-matches: id out of range");
      if (i && (m[i - 1].read_id > m[i].read_id)) throw std::runtime_error("--format-matches: records must be ordered by read id");
    }
  } else if (nd == 1) {
    m.swap(dm[0]);
  } else {
    size_t n = 0;
    for (auto& v : dm) n += v.size();
    m.reserve(n);
    for (auto& v : dm) {
      m.insert(m.end(), v.begin(), v.end());
      std::vector<msc_match>().swap(v);
    }
  }
  const double t_gpu = since(t_g0);

  // ---- outputs ------------------------------------------------------------------------------------
  const auto t_o0 = std::chrono::steady_clock::now();
  Outputs o;
  o.reads = &reads;
  o.targets = &targets;
  o.threads = opt.threads;
  // matches.txt.sz: whole-line sorted like `sort -u` leaves it (cmd/muscato/main.go:453-463)
  write_match_lines(o, m.data(), m.size(), false, cfg.TempDir + "/matches.txt.sz", true);
  if (keep_copy) write_match_lines(o, m.data(), m.size(), false, cfg.TempDir + "/msc_b200_matches.txt.sz", true);
  if (opt.epilogue) {
    // sortByGeneId + joinGeneNames + `sort -k1` + joinReadNames (cmd/muscato/main.go:507-676)
    GeneIds genes;
    genes.text = szio::read_text(cfg.GeneIdFileName);
    parse_gene_ids(genes);
    if (genes.name.size() < targets.n()) throw std::runtime_error("gene id file is shorter than the target file");
    o.genes = &genes;
    write_match_lines(o, m.data(), m.size(), true, cfg.ResultsFileName, false);
    // muscato_nonmatch (cmd/muscato_nonmatch/main.go:95-113) with an exact matched set: the device's own list
    // (msc_fetch_nonmatch) when every read saw the whole database in one call, else from the final matches
    std::vector<uint32_t> ids;
    bool all_valid = true;
    for (char v : nm_valid) all_valid = all_valid && v;
    if (all_valid && nd > 0) {
      for (auto& v : dn) ids.insert(ids.end(), v.begin(), v.end());
    } else {
      std::vector<char> matched(reads.n(), 0);
      for (const msc_match& x : m) matched[x.read_id] = 1;
      for (uint64_t i = 0; i < reads.n(); i++)
        if (!matched[i]) ids.push_back((uint32_t)i);
    }
    write_nonmatch(o, ids.data(), ids.size(), nonmatch_name(cfg.ResultsFileName));
  }
  const double t_out = since(t_o0);
  fprintf(stderr, "muscato_b200_hotpath: %llu reads, %llu target bases, %llu candidates, %llu pairs, %llu matches\n",
          (unsigned long long)reads.n(), (unsigned long long)targets.bases(), (unsigned long long)tot.candidates,
          (unsigned long long)tot.pairs, (unsigned long long)m.size());
  // JSON run report (bases screened, candidates, confirmed, stage seconds)
  const std::string rep = !opt.report.empty() ? opt.report : (!cfg.LogDir.empty() ? cfg.LogDir + "/muscato_b200_hotpath.json" : std::string());
  if (!rep.empty()) {
    if (FILE* f = fopen(rep.c_str(), "w")) {
      fprintf(f,
              "{\"reads\": %llu, \"reads_sorted_unique\": %s, \"targets\": %llu, \"target_bases\": %llu, \"target_cache\": \"%s\", "
              "\"devices\": %zu, \"tiles\": %llu, \"target_ranges\": %zu, \"candidates\": %llu, \"pairs\": %llu, \"matches\": %llu, "
              "\"h2d_bytes\": %llu, \"d2h_bytes\": %llu, \"scan_kernel_ms\": %.3f, "
              "\"seconds\": {\"targets\": %.3f, \"reads\": %.3f, \"gpu_path\": %.3f, \"outputs\": %.3f, \"total\": %.3f}}\n",
              (unsigned long long)reads.n(), reads.sorted_unique ? "true" : "false", (unsigned long long)targets.n(),
              (unsigned long long)targets.bases(), from_cache ? "hit" : (write_cache ? "written" : "off"), nd,
              (unsigned long long)tot.tiles, ranges.size(), (unsigned long long)tot.candidates, (unsigned long long)tot.pairs,
              (unsigned long long)m.size(), (unsigned long long)tot.h2d, (unsigned long long)tot.d2h, tot.ms_scan, t_targets, t_reads,
              t_gpu, t_out, since(t_start));
      fclose(f);
    }
  }
  return 0;
}

void write_empty_sz(const std::string& path) { szio::write_file(path, std::string(), true); }

std::string base_name(const char* p) {
  std::string s(p);
  const size_t sl = s.rfind('/');
  return sl == std::string::npos ? s : s.substr(sl + 1);
}

struct FileLock {
  int fd = -1;
  explicit FileLock(const std::string& path) {
    fd = open(path.c_str(), O_CREAT | O_RDWR, 0644);
    if (fd < 0) throw std::runtime_error("cannot open lock file " + path);
    if (flock(fd, LOCK_EX) != 0) throw std::runtime_error("flock failed on " + path);
  }
  ~FileLock() {
    if (fd >= 0) {
      flock(fd, LOCK_UN);
      close(fd);
    }
  }
};

bool file_exists(const std::string& p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0;
}

// muscato_screen <config.json> [tmpdir]  (cmd/muscato/main.go:310; cmd/muscato_screen/main.go:517-528)
int shim_screen(int argc, char** argv) {
  if (argc != 2 && argc != 3) {
    fprintf(stderr, "muscato_screen: wrong number of arguments\n");
    return 1;
  }
  Cfg cfg = read_cfg(argv[1]);
  if (argc == 3) cfg.TempDir = argv[2];
  Options opt;
  opt.epilogue = false;
  if (const char* d = getenv("MSC_DEVICE")) opt.devices = {atoi(d)};
  FileLock lk(cfg.TempDir + "/msc_b200.lock");
  const int rc = run_hotpath(cfg, opt, true);
  for (size_t k = 0; k < cfg.Windows.size(); k++) write_empty_sz(cfg.TempDir + "/bmatch_" + std::to_string(k) + ".txt.sz");
  return rc;
}

// muscato_confirm <config.json> <k>  (cmd/muscato/main.go:402; cmd/muscato_confirm/main.go:275-293)
int shim_confirm(int argc, char** argv) {
  if (argc != 3 && argc != 4) {
    fprintf(stderr, "muscato_confirm: wrong number of arguments\n");
    return 1;
  }
  Cfg cfg = read_cfg(argv[1]);
  if (argc == 4) cfg.TempDir = argv[3];
  const int k = atoi(argv[2]);
  if (k < 0 || k >= (int)cfg.Windows.size()) throw std::runtime_error("window index out of range");
  const std::string stored = cfg.TempDir + "/msc_b200_matches.txt.sz";
  {
    // concurrent callers (one per window, cmd/muscato/main.go:391-420) serialise here; the first one
    // runs the GPU path when muscato_screen was not ours
    FileLock lk(cfg.TempDir + "/msc_b200.lock");
    if (!file_exists(stored)) {
      Options opt;
      opt.epilogue = false;
      if (const char* d = getenv("MSC_DEVICE")) opt.devices = {atoi(d)};
      const int rc = run_hotpath(cfg, opt, true);
      if (rc) return rc;
    }
  }
  const std::string out = cfg.TempDir + "/rmatch_" + std::to_string(k) + ".txt.sz";
  if (k == 0) {
    const std::string raw = szio::read_all(stored);
    szio::write_file(out, raw, false);  // already framed
  } else {
    write_empty_sz(out);
  }
  return 0;
}

// muscato_combine_windows <config.json>: stdin (sorted, de-duplicated rmatch lines) -> stdout; per read
// keep the lines with nx <= best + MMTol in input order (cmd/muscato_combine_windows/main.go:36-60, :94-143).
int shim_combine_windows(int argc, char** argv) {
  if (argc != 2) {
    fprintf(stderr, "muscato_combine_windows: wrong number of arguments\n");
    return 1;
  }
  const Cfg cfg = read_cfg(argv[1]);
  std::string text;
  {
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, stdin)) > 0) text.append(buf, n);
  }
  std::vector<std::pair<size_t, size_t>> grp;  // (offset, length) of the current read's lines
  std::vector<int> nxs;
  std::string out;
  auto flush = [&]() {
    if (grp.empty()) return;
    int best = nxs[0];
    for (int v : nxs) best = std::min(best, v);
    for (size_t i = 0; i < grp.size(); i++)
      if (nxs[i] <= best + cfg.MMTol) {
        out.append(text, grp[i].first, grp[i].second);
        out += '\n';
      }
    grp.clear();
    nxs.clear();
    if (out.size() > (1u << 20)) {
      fwrite(out.data(), 1, out.size(), stdout);
      out.clear();
    }
  };
  size_t i = 0, cur_off = 0, cur_len = 0;
  bool have = false;
  while (i < text.size()) {
    size_t j = text.find('\n', i);
    if (j == std::string::npos) j = text.size();
    const size_t n = j - i;
    if (n) {
      // strings.Fields: whitespace-separated; field 0 = read, field 3 = nx
      size_t f[5], fl[5];
      int nf = 0;
      size_t p = i;
      while (p < j && nf < 5) {
        while (p < j && is_ws(text[p])) p++;
        if (p >= j) break;
        f[nf] = p;
        while (p < j && !is_ws(text[p])) p++;
        fl[nf] = p - f[nf];
        nf++;
      }
      if (nf >= 4) {
        if (!have || fl[0] != cur_len || memcmp(text.data() + f[0], text.data() + cur_off, cur_len) != 0) {
          flush();
          cur_off = f[0];
          cur_len = fl[0];
          have = true;
        }
        grp.emplace_back(i, n);
        nxs.push_back(atoi(std::string(text, f[3], fl[3]).c_str()));
      }
    }
    i = j + 1;
  }
  flush();
  fwrite(out.data(), 1, out.size(), stdout);
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  try {
    const std::string tool = base_name(argv[0]);
    if (tool == "muscato_screen") return shim_screen(argc, argv);
    if (tool == "muscato_confirm") return shim_confirm(argc, argv);
    if (tool == "muscato_combine_windows") return shim_combine_windows(argc, argv);
    if (argc < 2) {
      fprintf(stderr,
              "usage: %s <config.json> [--device N | --devices a,b,..] [--from-fastq] [--no-epilogue] [--no-target-cache]\n"
              "          [--threads N] [--max-items N] [--max-bases N] [--report FILE]\n",
              argv[0]);
      return 1;
    }
    // sztool-equivalent helpers (the reference shells out to `sztool -d f` / `sztool -c - f`,
    // cmd/muscato/main.go:255, :274): `--sz-cat f [threads]` and `--sz-pack in out` (in = "-" reads stdin).
    if (!strcmp(argv[1], "--sz-cat")) {
      if (argc < 3) throw std::runtime_error("--sz-cat needs a file");
      const std::string raw = szio::read_all(argv[2]);
      const std::string txt = szio::is_framed(raw) ? szio::decompress(raw, argc > 3 ? (unsigned)atoi(argv[3]) : 0u) : raw;
      fwrite(txt.data(), 1, txt.size(), stdout);
      return 0;
    }
    if (!strcmp(argv[1], "--sz-pack") && argc == 4) {
      std::string data;
      if (!strcmp(argv[2], "-")) {
        char buf[1 << 16];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, stdin)) > 0) data.append(buf, n);
      } else {
        data = szio::read_all(argv[2]);
      }
      szio::write_file(argv[3], data, true);
      return 0;
    }
    Options opt;
    for (int a = 2; a < argc; a++) {
      if (!strcmp(argv[a], "--device") && a + 1 < argc) opt.devices = {atoi(argv[++a])};
      else if (!strcmp(argv[a], "--devices") && a + 1 < argc) {
        opt.devices.clear();
        const std::string s = argv[++a];
        size_t i = 0;
        while (i <= s.size()) {
          size_t j = s.find(',', i);
          if (j == std::string::npos) j = s.size();
          if (j > i) opt.devices.push_back(atoi(s.substr(i, j - i).c_str()));
          i = j + 1;
        }
        if (opt.devices.empty()) throw std::runtime_error("--devices: empty list");
      } else if (!strcmp(argv[a], "--from-fastq")) opt.from_fastq = true;
      else if (!strcmp(argv[a], "--no-epilogue")) opt.epilogue = false;
      else if (!strcmp(argv[a], "--no-target-cache")) opt.target_cache = false;
      else if (!strcmp(argv[a], "--threads") && a + 1 < argc) opt.threads = (unsigned)std::max(1, atoi(argv[++a]));
      else if (!strcmp(argv[a], "--max-items") && a + 1 < argc) opt.max_items = strtoull(argv[++a], nullptr, 10);
      else if (!strcmp(argv[a], "--max-bases") && a + 1 < argc) opt.max_bases = strtoull(argv[++a], nullptr, 10);
      else if (!strcmp(argv[a], "--report") && a + 1 < argc) opt.report = argv[++a];
      else if (!strcmp(argv[a], "--format-matches") && a + 1 < argc) {
        opt.matches_file = argv[++a];
        opt.target_cache = false;
      }
      else throw std::runtime_error(std::string("unknown argument ") + argv[a]);
    }
    const Cfg cfg = read_cfg(argv[1]);
    return run_hotpath(cfg, opt, false);
  } catch (const std::exception& e) {
    fprintf(stderr, "muscato_b200_hotpath: %s\n", e.what());
    return 2;
  }
}
