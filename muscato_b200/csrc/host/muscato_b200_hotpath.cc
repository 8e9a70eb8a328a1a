// muscato_b200_hotpath -- stage-compatible executable for the hot path.
//
//   muscato_b200_hotpath <LogDir/config.json> [--device N] [--from-fastq] [--no-epilogue]
//
// Drop-in for steps 5-12 of the reference driver (cmd/muscato/main.go:1029-1051: screen,
// sortBloom, confirm, combineWindows, sortByGeneId, joinGeneNames, joinReadNames,
// writeNonMatch) with the reference's own contracts: the same config.json (utils/config.go),
// TempDir/reads_sorted.txt.sz, Config.GeneFileName, Config.GeneIdFileName in; 
// TempDir/matches.txt.sz, Config.ResultsFileName and the non-match fastq out.  All matching is
// done by libmuscato_b200.so on the GPU through the C ABI (include/muscato_b200.h); this file
// only parses and formats text.  Exit status != 0 on any failure, like the reference stages
// (log.Fatal / panic), upon which the driver panics (cmd/muscato/main.go:313-315).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/muscato_b200.h"
#include "szio.hpp"

namespace {

struct Cfg {
  std::string ReadFileName, GeneFileName, GeneIdFileName, ResultsFileName, TempDir, MatchMode;
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

struct Lines {
  std::string text;
  std::vector<std::pair<size_t, size_t>> ln;  // (offset, length), '\n' and one trailing '\r' stripped
  void split() {
    size_t i = 0;
    while (i < text.size()) {
      size_t j = text.find('\n', i);
      if (j == std::string::npos) j = text.size();
      size_t n = j - i;
      if (n && text[i + n - 1] == '\r') n--;
      ln.emplace_back(i, n);
      i = j + 1;
    }
  }
};

inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\v' || c == '\f' || c == '\r' || c == '\n'; }

struct Reads {
  std::vector<std::string> seq, count, names;
};

// reads_sorted.txt.sz: `seq \t count \t names` (cmd/muscato_uniqify/main.go:89-110); the screen and
// window_reads take bytes.Fields(line)[0] (cmd/muscato_screen/main.go:172).
Reads parse_reads_sorted(const std::string& text) {
  Lines L;
  L.text = text;
  L.split();
  Reads r;
  for (auto& pr : L.ln) {
    if (pr.second == 0) continue;
    const char* p = L.text.data() + pr.first;
    const size_t n = pr.second;
    size_t t1 = 0;
    while (t1 < n && p[t1] != '\t') t1++;
    size_t t2 = t1 + 1;
    while (t2 < n && p[t2] != '\t') t2++;
    size_t a = 0;
    while (a < t1 && is_ws(p[a])) a++;
    size_t b = a;
    while (b < t1 && !is_ws(p[b])) b++;
    r.seq.emplace_back(p + a, b - a);
    r.count.emplace_back(t1 < n ? std::string(p + t1 + 1, std::min(t2, n) - t1 - 1) : std::string());
    r.names.emplace_back(t2 < n ? std::string(p + t2 + 1, n - t2 - 1) : std::string());
  }
  return r;
}

// prepReads = muscato_prep_reads | sort | muscato_uniqify (cmd/muscato/main.go:152-221) with the
// sort and the collapse of equal sequences done on the GPU (msc_prep_reads).  The host parses the
// fastq records (utils/fastq.go:35-61), hands the raw sequences over, and joins counts and names
// from the grouping the library returns: inside a group the `seq\tname` lines are in bytewise
// name order, names longer than 1000 bytes are cut to 995 + "..." before sorting
// (cmd/muscato_prep_reads/main.go:76-79), the name column is the text before its first tab and
// the joined names are cut to 996 + "..." (cmd/muscato_uniqify/main.go:89-111).
Reads prep_reads_device(msc_ctx* ctx, const std::string& fastq, int min_len, int max_len) {
  Lines L;
  L.text = fastq;
  L.split();
  std::vector<std::string> names;
  std::string raw;
  std::vector<uint64_t> offs(1, 0);
  for (size_t i = 0; i + 4 <= L.ln.size(); i += 4) {
    names.emplace_back(L.text, L.ln[i].first, L.ln[i].second);
    raw.append(L.text, L.ln[i + 1].first, L.ln[i + 1].second);
    offs.push_back(raw.size());
  }
  const uint64_t n_raw = names.size();
  uint64_t kept = 0, uniq = 0;
  if (msc_prep_reads(ctx, reinterpret_cast<const uint8_t*>(raw.data()), offs.data(), n_raw, min_len, &kept, &uniq) != MSC_OK)
    throw std::runtime_error(std::string("msc_prep_reads: ") + msc_last_error(ctx));
  std::vector<uint32_t> perm(kept + 1), gs(uniq + 1);
  if (msc_fetch_read_groups(ctx, perm.data(), gs.data()) != MSC_OK)
    throw std::runtime_error(std::string("msc_fetch_read_groups: ") + msc_last_error(ctx));
  Reads r;
  std::vector<std::string> nm;
  for (uint64_t u = 0; u < uniq; u++) {
    nm.clear();
    for (uint32_t j = gs[u]; j < gs[u + 1]; j++) {
      std::string n = names[perm[j]];
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
    r.seq.push_back(seq);
    r.count.push_back(std::to_string(nm.size()));
    r.names.push_back(na);
  }
  return r;
}

std::vector<std::string> parse_targets(const std::string& text) {
  Lines L;
  L.text = text;
  L.split();
  std::vector<std::string> out;
  for (auto& pr : L.ln) {
    const char* p = L.text.data() + pr.first;
    size_t n = 0;
    while (n < pr.second && p[n] != '\t') n++;  // toks[0], cmd/muscato_screen/main.go:448-449
    out.emplace_back(p, n);
  }
  return out;
}

void parse_gene_ids(const std::string& text, std::vector<std::string>& names, std::vector<std::string>& lens) {
  Lines L;
  L.text = text;
  L.split();
  for (auto& pr : L.ln) {
    if (pr.second == 0) continue;
    const std::string line(L.text, pr.first, pr.second);
    const size_t t1 = line.find('\t');
    const size_t t2 = t1 == std::string::npos ? std::string::npos : line.find('\t', t1 + 1);
    names.push_back(t1 == std::string::npos ? std::string() : line.substr(t1 + 1, t2 == std::string::npos ? std::string::npos : t2 - t1 - 1));
    lens.push_back(t2 == std::string::npos ? std::string() : line.substr(t2 + 1));
  }
}

void concat(const std::vector<std::string>& v, std::string& all, std::vector<uint64_t>& offs) {
  offs.assign(v.size() + 1, 0);
  size_t tot = 0;
  for (size_t i = 0; i < v.size(); i++) { tot += v[i].size(); offs[i + 1] = tot; }
  all.clear();
  all.reserve(tot);
  for (auto& s : v) all += s;
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

}  // namespace

int main(int argc, char** argv) {
  try {
    if (argc < 2) {
      fprintf(stderr, "usage: %s <config.json> [--device N] [--from-fastq] [--no-epilogue]\n", argv[0]);
      return 1;
    }
    // sztool-equivalent helpers (the reference shells out to `sztool -d f` / `sztool -c - f`,
    // cmd/muscato/main.go:255, :274): `--sz-cat f` and `--sz-pack in out`.
    if (!strcmp(argv[1], "--sz-cat") && argc == 3) {
      const std::string t = szio::read_text(argv[2]);
      fwrite(t.data(), 1, t.size(), stdout);
      return 0;
    }
    if (!strcmp(argv[1], "--sz-pack") && argc == 4) {
      szio::write_file(argv[3], szio::read_all(argv[2]), true);
      return 0;
    }
    if (!strcmp(argv[1], "--sz-cat")) {  // decode a .sz file to stdout (no GPU involved): muscato_b200_hotpath --sz-cat FILE [threads]
      if (argc < 3) throw std::runtime_error("--sz-cat needs a file");
      const std::string raw = szio::read_all(argv[2]);
      const std::string txt = szio::is_framed(raw) ? szio::decompress(raw, argc > 3 ? (unsigned)atoi(argv[3]) : 0u) : raw;
      fwrite(txt.data(), 1, txt.size(), stdout);
      return 0;
    }
    int device = 0;
    bool from_fastq = false, epilogue = true;
    for (int a = 2; a < argc; a++) {
      if (!strcmp(argv[a], "--device") && a + 1 < argc) device = atoi(argv[++a]);
      else if (!strcmp(argv[a], "--from-fastq")) from_fastq = true;
      else if (!strcmp(argv[a], "--no-epilogue")) epilogue = false;
      else throw std::runtime_error(std::string("unknown argument ") + argv[a]);
    }
    const Cfg cfg = read_cfg(argv[1]);

    const std::vector<std::string> targets = parse_targets(szio::read_text(cfg.GeneFileName));

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
    char err[512] = {0};
    msc_ctx* ctx = msc_create(&mc, err, sizeof err);
    if (!ctx) throw std::runtime_error(std::string("msc_create: ") + err);
    auto check = [&](int rc, const char* what) {
      if (rc != MSC_OK) {
        const std::string m = std::string(what) + ": " + msc_last_error(ctx);
        msc_destroy(ctx);
        throw std::runtime_error(m);
      }
    };

    Reads reads;
    const std::string rs_path = cfg.TempDir + "/reads_sorted.txt.sz";
    if (from_fastq) {
      // prepReads with the sort / uniqify on the device; the unique reads are installed as the
      // read set by msc_prep_reads itself
      try {
        reads = prep_reads_device(ctx, szio::read_all(cfg.ReadFileName), cfg.MinReadLength, cfg.MaxReadLength);
      } catch (...) {
        msc_destroy(ctx);
        throw;
      }
      std::string txt;
      for (size_t i = 0; i < reads.seq.size(); i++) txt += reads.seq[i] + "\t" + reads.count[i] + "\t" + reads.names[i] + "\n";
      szio::write_file(rs_path, txt, true);
    } else {
      reads = parse_reads_sorted(szio::read_text(rs_path));
      std::string all;
      std::vector<uint64_t> offs;
      concat(reads.seq, all, offs);
      check(msc_set_reads(ctx, reinterpret_cast<const uint8_t*>(all.data()), offs.data(), reads.seq.size()), "msc_set_reads");
    }
    {
      std::string all;
      std::vector<uint64_t> offs;
      concat(targets, all, offs);
      check(msc_set_targets(ctx, reinterpret_cast<const uint8_t*>(all.data()), offs.data(), targets.size()), "msc_set_targets");
    }
    check(msc_run(ctx), "msc_run");
    msc_match* m = nullptr;
    uint64_t n = 0;
    check(msc_fetch_matches(ctx, &m, &n), "msc_fetch_matches");
    msc_stats st;
    msc_get_stats(ctx, &st);

    // matches.txt.sz: read \t target[pos:pos+L] \t pos \t nx \t %011d(gene) (cmd/muscato_confirm/main.go:221-230),
    // whole-line sorted like `sort -u` leaves it (cmd/muscato/main.go:453-463).
    std::vector<std::string> lines(n);
    char tail[64];
    for (uint64_t i = 0; i < n; i++) {
      const std::string& r = reads.seq[m[i].read_id];
      snprintf(tail, sizeof tail, "\t%u\t%u\t%011u", m[i].pos, m[i].nx, m[i].gene_id);
      lines[i] = r + "\t" + targets[m[i].gene_id].substr(m[i].pos, r.size()) + tail;
    }
    std::sort(lines.begin(), lines.end());
    {
      std::string txt;
      for (auto& l : lines) { txt += l; txt += '\n'; }
      szio::write_file(cfg.TempDir + "/matches.txt.sz", txt, true);
    }
    if (epilogue) {
      // sortByGeneId + joinGeneNames + `sort -k1` + joinReadNames (cmd/muscato/main.go:507-676)
      std::vector<std::string> gnames, glens;
      parse_gene_ids(szio::read_text(cfg.GeneIdFileName), gnames, glens);
      std::vector<std::pair<std::string, uint32_t>> rows(n);
      for (uint64_t i = 0; i < n; i++) {
        const std::string& r = reads.seq[m[i].read_id];
        const uint32_t g = m[i].gene_id;
        if (g >= gnames.size()) throw std::runtime_error("gene id file is shorter than the target file");
        snprintf(tail, sizeof tail, "\t%u\t%u\t", m[i].pos, m[i].nx);
        rows[i] = {r + "\t" + targets[g].substr(m[i].pos, r.size()) + tail + gnames[g] + "\t" + glens[g], m[i].read_id};
      }
      std::sort(rows.begin(), rows.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
      std::string txt;
      for (auto& row : rows) txt += row.first + "\t" + reads.count[row.second] + "\t" + reads.names[row.second] + "\n";
      szio::write_file(cfg.ResultsFileName, txt, false);
      // muscato_nonmatch (cmd/muscato_nonmatch/main.go:95-113) with an exact matched set
      std::vector<char> matched(reads.seq.size(), 0);
      for (uint64_t i = 0; i < n; i++) matched[m[i].read_id] = 1;
      std::string fq;
      for (size_t i = 0; i < reads.seq.size(); i++) {
        if (matched[i]) continue;
        const std::string& nm = reads.names[i];
        size_t a = 0;
        while (a < nm.size() && is_ws(nm[a])) a++;
        size_t b = a;
        while (b < nm.size() && !is_ws(nm[b])) b++;
        fq += nm.substr(a, b - a) + "#" + reads.count[i] + "\n" + reads.seq[i] + "\n+\n" + std::string(reads.seq[i].size(), '!') + "\n";
      }
      szio::write_file(nonmatch_name(cfg.ResultsFileName), fq, false);
    }
    fprintf(stderr, "muscato_b200_hotpath: %llu reads, %llu target bases, %llu candidates, %llu pairs, %llu matches\n",
            (unsigned long long)st.n_reads, (unsigned long long)st.target_bases, (unsigned long long)st.n_candidates,
            (unsigned long long)st.n_pairs, (unsigned long long)st.n_matches);
    msc_free(m);
    msc_destroy(ctx);
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "muscato_b200_hotpath: %s\n", e.what());
    return 2;
  }
}
