// muscato_oracle.cc -- TEST INFRASTRUCTURE ONLY (not product code).
//
// CPU restatement ("port") of the kshedden/muscato pipeline, stage by stage,
// used (a) as the parity oracle for the CUDA hot path in tests/, smoke() and
// (b) as the `cpu_baseline` / `--impl reference` arm of bench.py.  Nothing in
// muscato_b200/ may include, link or execute this file.
//
// The reference is Go (no Go toolchain in this image, six un-vendored modules
// missing), so it cannot be compiled here; this file restates its algorithm
// in C++17 and is pinned against the reference's own golden fixtures
// (tests/golden/muscato/00..04, tests/golden/prep_targets/00..07).  GNU
// sort / join / cut are invoked exactly where the reference's driver invokes
// them (LC_ALL=C).  All citations are file:line under /root/reference.
//
// Deliberate, documented deviations from the Go code:
//  * intermediates are plain text files instead of Snappy-framed ".sz" files
//    (container format only; content is identical text);
//  * Q8: an empty smatch_k / win_k_sorted stream yields empty rmatch_k instead
//    of the reference's index-out-of-range panic (cmd/muscato_confirm/main.go:382);
//  * Q9: muscato_combine_filter's lossy Bloom pre-dedup is replaced by plain
//    concatenation (the exact `sort -u` that follows does the dedup);
//  * Q10: muscato_nonmatch's Bloom membership is replaced by an exact set;
//  * third-party buzhash32 (github.com/chmduquesne/rollinghash, un-pinned) is
//    restated from its published algorithm (cyclic polynomial: sum = rotl(sum,1)
//    ^ rotl(T[out], n mod 32) ^ T[in]); its byte tables come from a fixed-seed
//    splitmix64 instead of Go's unseeded math/rand (they never reach results:
//    the merge join in confirm discards Bloom false positives).
//
// Sub-commands (see main()):
//   prep_targets [-rev] <in.txt|in.fasta> <seq_out> <ids_out>
//   pipeline <config.json>     whole muscato run (steps 1..12 of cmd/muscato/main.go:1005-1058)
//   upstream <config.json>     prepReads + windowReads + sortWindows only
//   windows  <config.json>     windowReads + sortWindows on an existing TempDir/reads_sorted.txt
//   hotpath  <config.json>     screen + sortBloom + confirm + combineWindows only (timed)
//   epilogue <config.json>     sortByGeneId + joinGeneNames + joinReadNames + nonmatch
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

#include <sys/wait.h>
#include <unistd.h>

namespace {

[[noreturn]] void die(const std::string& msg) {
  fprintf(stderr, "muscato_oracle: %s\n", msg.c_str());
  exit(2);
}

double now_s() {
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------------------
// Config: mirrors utils.Config (utils/config.go:10-101) field for field.
// ---------------------------------------------------------------------------
struct Config {
  std::string ReadFileName, GeneFileName, GeneIdFileName, ResultsFileName;
  std::vector<int> Windows;
  int WindowWidth = 0;
  uint64_t BloomSize = 0;
  int NumHash = 0;
  double PMatch = 0;
  int MinDinuc = 0;
  std::string TempDir, LogDir;
  int MinReadLength = 0, MaxReadLength = 0, MaxMatches = 0, MaxConfirmProcs = 0, MMTol = 0;
  std::string MatchMode;
  int SortPar = 0;
  std::string SortTemp, SortMem;
  bool NoCleanTemp = false;
  int Threads = 0;  // oracle-only knob (0 = hardware concurrency)
};

// Minimal JSON reader for the flat objects utils.ReadConfig decodes (utils/config.go:103-117).
struct JsonCur {
  const std::string& s;
  size_t i = 0;
  explicit JsonCur(const std::string& str) : s(str) {}
  void ws() { while (i < s.size() && strchr(" \t\r\n", s[i])) i++; }
  bool eat(char c) { ws(); if (i < s.size() && s[i] == c) { i++; return true; } return false; }
  std::string str() {
    ws();
    if (s[i] != '"') die("json: expected string");
    i++;
    std::string out;
    while (i < s.size() && s[i] != '"') {
      if (s[i] == '\\' && i + 1 < s.size()) {
        char c = s[++i];
        switch (c) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u': { out += (char)strtol(s.substr(i + 1, 4).c_str(), nullptr, 16); i += 4; break; }
          default: out += c;
        }
        i++;
      } else {
        out += s[i++];
      }
    }
    i++;
    return out;
  }
  std::string scalar() {  // number / true / false / null token
    ws();
    size_t j = i;
    while (i < s.size() && !strchr(",]} \t\r\n", s[i])) i++;
    return s.substr(j, i - j);
  }
};

Config read_config(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) die("cannot open config " + path);
  std::string txt;
  char buf[4096];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) txt.append(buf, n);
  fclose(f);
  Config c;
  JsonCur j(txt);
  if (!j.eat('{')) die("json: expected object");
  while (true) {
    j.ws();
    if (j.eat('}')) break;
    std::string key = j.str();
    if (!j.eat(':')) die("json: expected ':'");
    j.ws();
    if (txt[j.i] == '"') {
      std::string v = j.str();
      if (key == "ReadFileName") c.ReadFileName = v;
      else if (key == "GeneFileName") c.GeneFileName = v;
      else if (key == "GeneIdFileName") c.GeneIdFileName = v;
      else if (key == "ResultsFileName") c.ResultsFileName = v;
      else if (key == "TempDir") c.TempDir = v;
      else if (key == "LogDir") c.LogDir = v;
      else if (key == "MatchMode") c.MatchMode = v;
      else if (key == "SortTemp") c.SortTemp = v;
      else if (key == "SortMem") c.SortMem = v;
    } else if (txt[j.i] == '[') {
      j.eat('[');
      std::vector<int> arr;
      while (!j.eat(']')) {
        arr.push_back(atoi(j.scalar().c_str()));
        j.eat(',');
      }
      if (key == "Windows") c.Windows = arr;
    } else {
      std::string v = j.scalar();
      if (key == "WindowWidth") c.WindowWidth = atoi(v.c_str());
      else if (key == "BloomSize") c.BloomSize = strtoull(v.c_str(), nullptr, 10);
      else if (key == "NumHash") c.NumHash = atoi(v.c_str());
      else if (key == "PMatch") c.PMatch = strtod(v.c_str(), nullptr);
      else if (key == "MinDinuc") c.MinDinuc = atoi(v.c_str());
      else if (key == "MinReadLength") c.MinReadLength = atoi(v.c_str());
      else if (key == "MaxReadLength") c.MaxReadLength = atoi(v.c_str());
      else if (key == "MaxMatches") c.MaxMatches = atoi(v.c_str());
      else if (key == "MaxConfirmProcs") c.MaxConfirmProcs = atoi(v.c_str());
      else if (key == "MMTol") c.MMTol = atoi(v.c_str());
      else if (key == "SortPar") c.SortPar = atoi(v.c_str());
      else if (key == "NoCleanTemp") c.NoCleanTemp = (v == "true");
      else if (key == "Threads") c.Threads = atoi(v.c_str());
    }
    j.eat(',');
  }
  // Defaults of checkArgs (cmd/muscato/main.go:859-903).
  if (c.BloomSize == 0) c.BloomSize = 4000000000ull;
  if (c.NumHash == 0) c.NumHash = 20;
  if (c.PMatch == 0) c.PMatch = 1;
  if (c.MaxMatches == 0) c.MaxMatches = 1000000;
  if (c.MaxConfirmProcs == 0) c.MaxConfirmProcs = 3;
  if (c.MatchMode.empty()) c.MatchMode = "best";
  if (c.SortPar == 0) c.SortPar = 8;
  if (c.SortMem.empty()) c.SortMem = "50%";
  if (c.Threads <= 0) {
    const char* e = getenv("MUSCATO_ORACLE_THREADS");
    c.Threads = e ? atoi(e) : (int)std::thread::hardware_concurrency();
    if (c.Threads <= 0) c.Threads = 1;
  }
  return c;
}

// ---------------------------------------------------------------------------
// Text helpers
// ---------------------------------------------------------------------------
// Line reader with bufio.Scanner(ScanLines) semantics: '\n' terminated, one
// trailing '\r' dropped, final unterminated line returned.
struct LineReader {
  FILE* f;
  char* buf = nullptr;
  size_t cap = 0;
  explicit LineReader(const std::string& path) {
    f = fopen(path.c_str(), "rb");
    if (!f) die("cannot open " + path);
  }
  ~LineReader() { if (f) fclose(f); free(buf); }
  bool next(std::string& line) {
    ssize_t n = getline(&buf, &cap, f);
    if (n < 0) return false;
    if (n > 0 && buf[n - 1] == '\n') n--;
    if (n > 0 && buf[n - 1] == '\r') n--;
    line.assign(buf, (size_t)n);
    return true;
  }
};

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; }

// bytes.Fields / strings.Fields on ASCII input.
std::vector<std::string> fields(const std::string& s) {
  std::vector<std::string> out;
  size_t i = 0, n = s.size();
  while (i < n) {
    while (i < n && is_space(s[i])) i++;
    if (i >= n) break;
    size_t j = i;
    while (j < n && !is_space(s[j])) j++;
    out.emplace_back(s, i, j - i);
    i = j;
  }
  return out;
}

std::vector<std::string> split_tab(const std::string& s) {
  std::vector<std::string> out;
  size_t i = 0;
  while (true) {
    size_t j = s.find('\t', i);
    if (j == std::string::npos) { out.emplace_back(s, i); break; }
    out.emplace_back(s, i, j - i);
    i = j + 1;
  }
  return out;
}

void sh(const std::string& cmd) {
  std::string full = "export LC_ALL=C; set -e; set -o pipefail; " + cmd;  // setupEnvs, cmd/muscato/main.go:906-912
  pid_t pid = fork();
  if (pid < 0) die("fork failed");
  if (pid == 0) {
    execl("/bin/bash", "bash", "-c", full.c_str(), (char*)nullptr);
    _exit(127);
  }
  int status = 0;
  if (waitpid(pid, &status, 0) < 0 || !WIFEXITED(status) || WEXITSTATUS(status) != 0)
    die("command failed: " + cmd);
}

std::string sort_args(const Config& c) {
  // sortmem = "-S <SortMem>", sortpar = "--parallel=<SortPar>" (cmd/muscato/main.go:893-903)
  std::string a = "-S " + c.SortMem + " --parallel=" + std::to_string(c.SortPar);
  if (!c.SortTemp.empty()) a += " --temporary-directory=" + c.SortTemp;
  return a;
}

std::string tmp(const Config& c, const std::string& name) { return c.TempDir + "/" + name; }

// subx: non A/T/G/C -> 'X' (cmd/muscato_prep_reads/main.go:33-44, cmd/muscato_prep_targets/main.go:69-80)
void subx(std::string& s) {
  for (auto& ch : s)
    if (ch != 'A' && ch != 'T' && ch != 'C' && ch != 'G') ch = 'X';
}

// revcomp (cmd/muscato_prep_targets/main.go:48-66): bytes other than A/T/G/C/X become 0x00.
std::string revcomp(const std::string& s) {
  std::string b(s.size(), '\0');
  size_t m = s.size() - 1;
  for (size_t i = 0; i < s.size(); i++) {
    switch (s[i]) {
      case 'A': b[m - i] = 'T'; break;
      case 'T': b[m - i] = 'A'; break;
      case 'G': b[m - i] = 'C'; break;
      case 'C': b[m - i] = 'G'; break;
      case 'X': b[m - i] = 'X'; break;
    }
  }
  return b;
}

// CountDinuc (utils/entropy.go:5-40): distinct adjacent pairs over {A,T,G,C,other}.
int count_dinuc(const char* seq, int n) {
  int wk[25] = {0};
  int last = 0, cnt = 0;
  for (int i = 0; i < n; i++) {
    int v;
    switch (seq[i]) {
      case 'A': v = 0; break;
      case 'T': v = 1; break;
      case 'G': v = 2; break;
      case 'C': v = 3; break;
      default: v = 4;
    }
    if (i > 0) {
      int k = 5 * last + v;
      if (wk[k] == 0) cnt++;
      wk[k]++;
    }
    last = v;
  }
  return cnt;
}

// ---------------------------------------------------------------------------
// muscato_prep_targets (cmd/muscato_prep_targets/main.go:82-213).  Plain-text in/out;
// the caller strips .gz/.sz containers.
// ---------------------------------------------------------------------------
void prep_targets(const std::string& in, const std::string& seqout, const std::string& idout, bool rev,
                  bool fasta) {
  LineReader rd(in);
  FILE* so = fopen(seqout.c_str(), "wb");
  FILE* io = fopen(idout.c_str(), "wb");
  if (!so || !io) die("prep_targets: cannot create outputs");
  std::string line;
  long lnum = 0;
  if (!fasta) {
    // processText :82-141
    while (rd.next(line)) {
      if (line.empty()) break;  // :94-96
      auto toks = split_tab(line);
      if (toks.size() != 2) { fclose(so); fclose(io); exit(0); }  // :99-103 (os.Exit(0))
      std::string nam = toks[0], seq = toks[1];
      subx(seq);
      fprintf(so, "%s\n", seq.c_str());
      if (rev) { std::string r = revcomp(seq); fwrite(r.data(), 1, r.size(), so); fputc('\n', so); }
      fprintf(io, "%011ld\t%s\t%zu\n", lnum, nam.c_str(), seq.size());
      lnum++;
      if (rev) { fprintf(io, "%011ld\t%s_r\t%zu\n", lnum, nam.c_str(), seq.size()); lnum++; }
    }
  } else {
    // processFasta :143-213 (note: the final record is NOT passed through subx, :204-212)
    std::string seqname, seq;
    auto flush = [&](bool r) {
      fwrite(seq.data(), 1, seq.size(), so);
      fputc('\n', so);
      fprintf(io, "%011ld\t%s%s\t%zu\n", lnum, seqname.c_str(), r ? "_r" : "", seq.size());
    };
    while (rd.next(line)) {
      if (line.empty()) die("prep_targets: empty fasta line (reference panics on line[0])");
      if (line[0] == '>') {
        if (!seq.empty()) {
          subx(seq);
          flush(false);
          lnum++;
          if (rev) { seq = revcomp(seq); flush(true); lnum++; }
        }
        seqname = line;
        seq.clear();
        continue;
      }
      seq += line;
    }
    if (!seq.empty()) {
      flush(false);
      lnum++;
      if (rev) { seq = revcomp(seq); flush(true); lnum++; }
    }
  }
  fclose(so);
  fclose(io);
}

// ---------------------------------------------------------------------------
// prepReads = muscato_prep_reads | sort | muscato_uniqify  (cmd/muscato/main.go:152-221)
// ---------------------------------------------------------------------------
void prep_reads(const Config& c) {
  // muscato_prep_reads source() (cmd/muscato_prep_reads/main.go:46-92); fastq reader utils/fastq.go:35-61
  {
    LineReader rd(c.ReadFileName);
    FILE* out = fopen(tmp(c, "reads_raw.txt").c_str(), "wb");
    if (!out) die("cannot create reads_raw.txt");
    std::string l[4];
    while (true) {
      bool ok = true;
      for (int j = 0; j < 4; j++)
        if (!rd.next(l[j])) { ok = false; break; }
      if (!ok) break;
      std::string name = l[0], seq = l[1];
      if ((int)seq.size() < c.MinReadLength) continue;  // :59-62
      subx(seq);                                        // :64-65
      if ((int)seq.size() > c.MaxReadLength) seq.resize(c.MaxReadLength);  // :67-69
      if (name.size() > 1000) name = name.substr(0, 995) + "...";          // :76-79
      fprintf(out, "%s\t%s\n", seq.c_str(), name.c_str());
    }
    fclose(out);
  }
  sh("sort " + sort_args(c) + " " + tmp(c, "reads_raw.txt") + " > " + tmp(c, "reads_rawsorted.txt"));
  // muscato_uniqify (cmd/muscato_uniqify/main.go:68-134)
  {
    LineReader rd(tmp(c, "reads_rawsorted.txt"));
    FILE* out = fopen(tmp(c, "reads_sorted.txt").c_str(), "wb");
    if (!out) die("cannot create reads_sorted.txt");
    std::string line;
    if (!rd.next(line)) die("muscato_uniqify: no input");  // :69-75
    auto toks = split_tab(line);
    std::string seq = toks[0];
    std::vector<std::string> names{toks.size() > 1 ? toks[1] : std::string()};
    auto printrow = [&]() {  // :89-111
      std::string na;
      for (size_t i = 0; i < names.size(); i++) { if (i) na += ';'; na += names[i]; }
      if (na.size() > 1000) na = na.substr(0, 996) + "...";
      fprintf(out, "%s\t%zu\t%s\n", seq.c_str(), names.size(), na.c_str());
    };
    while (rd.next(line)) {
      toks = split_tab(line);
      if (toks[0] != seq) { printrow(); seq = toks[0]; names.clear(); }
      names.push_back(toks.size() > 1 ? toks[1] : std::string());
    }
    printrow();
    fclose(out);
  }
}

// muscato_window_reads (cmd/muscato_window_reads/main.go:94-151)
void window_reads(const Config& c) {
  size_t nw = c.Windows.size();
  std::vector<FILE*> outs(nw);
  for (size_t k = 0; k < nw; k++) {
    outs[k] = fopen(tmp(c, "win_" + std::to_string(k) + ".txt").c_str(), "wb");
    if (!outs[k]) die("cannot create win file");
  }
  std::vector<long> nread(nw, 0);
  LineReader rd(tmp(c, "reads_sorted.txt"));
  std::string line;
  while (rd.next(line)) {
    auto f = fields(line);
    if (f.empty()) die("window_reads: blank line");
    const std::string& seq = f[0];
    for (size_t k = 0; k < nw; k++) {
      int q1 = c.Windows[k], q2 = q1 + c.WindowWidth;
      if ((int)seq.size() < q2) continue;  // :110-112
      nread[k]++;
      if (count_dinuc(seq.data() + q1, c.WindowWidth) < c.MinDinuc) continue;  // :115-118
      fwrite(seq.data() + q1, 1, c.WindowWidth, outs[k]);
      fputc('\t', outs[k]);
      fwrite(seq.data(), 1, q1, outs[k]);
      fputc('\t', outs[k]);
      fwrite(seq.data() + q2, 1, seq.size() - q2, outs[k]);
      fputc('\n', outs[k]);
    }
  }
  for (auto* f : outs) fclose(f);
  for (size_t k = 0; k < nw; k++)
    if (nread[k] == 0) {  // :143-151
      fprintf(stderr, "Window %zu produced no valid reads, exiting", k);
      exit(1);
    }
}

// sortWindows (cmd/muscato/main.go:237-304)
void sort_windows(const Config& c) {
  for (size_t k = 0; k < c.Windows.size(); k++)
    sh("sort " + sort_args(c) + " -k1 " + tmp(c, "win_" + std::to_string(k) + ".txt") + " > " +
       tmp(c, "win_" + std::to_string(k) + "_sorted.txt"));
}

// ---------------------------------------------------------------------------
// A tiny bounded thread pool standing in for "limit <- true; go f()" (screen :451-452, confirm :392-393)
// ---------------------------------------------------------------------------
class Pool {
 public:
  explicit Pool(int n) : cap_(4 * n + 4) {
    for (int i = 0; i < n; i++) th_.emplace_back([this] { run(); });
  }
  ~Pool() { finish(); }
  void submit(std::function<void()> fn) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_space_.wait(lk, [&] { return q_.size() < cap_; });
    q_.push(std::move(fn));
    cv_work_.notify_one();
  }
  void finish() {
    {
      std::unique_lock<std::mutex> lk(mu_);
      if (done_) return;
      done_ = true;
    }
    cv_work_.notify_all();
    for (auto& t : th_) t.join();
  }

 private:
  void run() {
    while (true) {
      std::function<void()> fn;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_work_.wait(lk, [&] { return done_ || !q_.empty(); });
        if (q_.empty()) return;
        fn = std::move(q_.front());
        q_.pop();
        cv_space_.notify_one();
      }
      fn();
    }
  }
  std::mutex mu_;
  std::condition_variable cv_work_, cv_space_;
  std::queue<std::function<void()>> q_;
  std::vector<std::thread> th_;
  size_t cap_;
  bool done_ = false;
};

// ---------------------------------------------------------------------------
// muscato_screen (cmd/muscato_screen/main.go)
// ---------------------------------------------------------------------------
struct Screen {
  const Config& c;
  int nh, W;
  std::vector<std::vector<uint32_t>> tables;  // genTables :86-101
  std::vector<uint64_t*> smp;                 // one bit array per window :549-552 (calloc: lazily zeroed pages, like Go's make)
  std::vector<FILE*> outs;                    // bmatch_k (harvest :369-403)
  std::vector<std::mutex> out_mu;
  std::atomic<uint64_t> bases{0}, recs{0};

  explicit Screen(const Config& cfg) : c(cfg), nh(cfg.NumHash), W(cfg.WindowWidth), out_mu(cfg.Windows.size()) {
    uint64_t st = 0x243F6A8885A308D3ull;  // fixed seed (reference: unseeded math/rand)
    auto next = [&]() {
      st += 0x9E3779B97F4A7C15ull;
      uint64_t z = st;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      return z ^ (z >> 31);
    };
    tables.assign(nh, std::vector<uint32_t>(256));
    for (int j = 0; j < nh; j++) {
      std::unordered_set<uint32_t> seen;
      for (int i = 0; i < 256; i++) {
        while (true) {
          uint32_t x = (uint32_t)next();
          if (seen.insert(x).second) { tables[j][i] = x; break; }
        }
      }
    }
    for (size_t k = 0; k < c.Windows.size(); k++) {
      uint64_t* p = (uint64_t*)calloc((c.BloomSize + 63) / 64 + 1, sizeof(uint64_t));
      if (!p) die("cannot allocate Bloom bit array");
      smp.push_back(p);
    }
  }
  ~Screen() { for (auto* p : smp) free(p); }

  static inline uint32_t rotl(uint32_t x, unsigned r) { r &= 31; return r ? (x << r) | (x >> (32 - r)) : x; }

  // buzhash32 Write over a fresh state.
  uint32_t hash_full(int j, const char* s, int n) const {
    uint32_t sum = 0;
    for (int i = 0; i < n; i++) sum = rotl(sum, 1) ^ tables[j][(uint8_t)s[i]];
    return sum;
  }

  void set_bit(size_t k, uint64_t x) {
    __atomic_fetch_or(&smp[k][x >> 6], 1ull << (x & 63), __ATOMIC_RELAXED);
  }
  bool get_bit(size_t k, uint64_t x) const { return (smp[k][x >> 6] >> (x & 63)) & 1; }

  // buildBloom :116-207.  The reference feeds one worker goroutine per window (:136-162); here the
  // read scan is done once per window thread, which sets the same bits.
  void build_bloom() {
    std::vector<std::thread> th;
    for (size_t k = 0; k < c.Windows.size(); k++) {
      th.emplace_back([this, k] {
        LineReader rd(tmp(c, "reads_sorted.txt"));
        std::string line;
        const int q1 = c.Windows[k], q2 = q1 + W;
        while (rd.next(line)) {
          size_t e = 0;
          while (e < line.size() && !is_space(line[e])) e++;  // bytes.Fields(line)[0] :172
          if (q2 > (int)e) continue;                                      // :177-179
          if (count_dinuc(line.data() + q1, W) < c.MinDinuc) continue;    // :183-185
          for (int j = 0; j < nh; j++)                                    // :150-159
            set_bit(k, (uint64_t)hash_full(j, line.data() + q1, W) % c.BloomSize);
        }
      });
    }
    for (auto& t : th) t.join();
  }

  // checkWin :220-253
  void check_win(const uint32_t* sums, uint64_t* iw, std::vector<int>& ix) const {
    for (int j = 0; j < nh; j++) iw[j] = (uint64_t)sums[j] % c.BloomSize;
    ix.clear();
    for (size_t k = 0; k < smp.size(); k++) {
      bool g = true;
      for (int j = 0; j < nh; j++)
        if (!get_bit(k, iw[j])) { g = false; break; }
      if (g) ix.push_back((int)k);
    }
  }

  void emit(std::vector<std::string>& local, int k, const char* mseq, int mlen, const char* left, int llen,
            const char* right, int rlen, long tnum, long pos) {
    // rec serialisation, harvest :389-400
    std::string& b = local[k];
    b.append(mseq, mlen);
    b += '\t';
    b.append(left, llen);
    b += '\t';
    b.append(right, rlen);
    char t[64];
    int n = snprintf(t, sizeof t, "\t%011ld\t%ld\n", tnum, pos);
    b.append(t, n);
    recs++;
    if (b.size() > (1u << 20)) flush(local, k);
  }
  void flush(std::vector<std::string>& local, int k) {
    if (local[k].empty()) return;
    std::lock_guard<std::mutex> lk(out_mu[k]);
    fwrite(local[k].data(), 1, local[k].size(), outs[k]);
    local[k].clear();
  }

  // processSeq :256-366
  void process_seq(const std::string& seq, long genenum, std::vector<std::string>& local) {
    int hlen = W, len = (int)seq.size();
    if (len < hlen) return;  // :268-271
    bases += (uint64_t)len;
    std::vector<uint32_t> sums(nh);
    std::vector<uint64_t> iw(nh);
    std::vector<int> ix;
    const unsigned nrot = (unsigned)hlen % 32;
    for (int j = 0; j < nh; j++) sums[j] = hash_full(j, seq.data(), hlen);  // :272-278
    check_win(sums.data(), iw.data(), ix);
    for (int i : ix) {  // :294-316
      int q1 = c.Windows[i];
      if (q1 != 0) continue;
      int q2 = q1 + W;
      int jz = 100 - q2;  // the literal 100 (Q1)
      if (jz > len) jz = len;
      if (jz < hlen) die("screen: slice bounds out of range at target position 0 (reference panics: 100-q2 < W)");
      emit(local, i, seq.data(), hlen, "", 0, seq.data() + hlen, jz - hlen, genenum, 0);
    }
    for (int j = hlen; j < len; j++) {  // :319-365
      uint8_t in = (uint8_t)seq[j], out = (uint8_t)seq[j - hlen];
      for (int h = 0; h < nh; h++)  // Roll: sum = rotl(sum,1) ^ rotl(T[out], n%32) ^ T[in]
        sums[h] = rotl(sums[h], 1) ^ rotl(tables[h][out], nrot) ^ tables[h][in];
      check_win(sums.data(), iw.data(), ix);
      for (int i : ix) {
        int q1 = c.Windows[i], q2 = q1 + W;
        if (j < q2 - 1) continue;  // :335-338
        int jx = j - hlen + 1, jy = j + 1;
        int jw = jx - q1;
        int jz = jy + c.MaxReadLength - q2;
        if (jz > len) jz = len;
        if (jw >= 0)
          emit(local, i, seq.data() + jx, jy - jx, seq.data() + jw, jx - jw, seq.data() + jy, std::max(0, jz - jy),
               genenum, jx);
      }
    }
  }

  // search :408-480
  void search() {
    size_t nw = c.Windows.size();
    outs.resize(nw);
    for (size_t k = 0; k < nw; k++) {
      outs[k] = fopen(tmp(c, "bmatch_" + std::to_string(k) + ".txt").c_str(), "wb");
      if (!outs[k]) die("cannot create bmatch file");
    }
    {
      // Targets are batched so that worker threads amortise the queue; gene# = 0-based line index (:440-452).
      Pool pool(c.Threads);
      LineReader rd(c.GeneFileName);
      std::string line;
      long i = 0;
      auto batch = std::make_shared<std::vector<std::pair<long, std::string>>>();
      size_t batch_bytes = 0;
      auto submit = [&]() {
        auto b = batch;
        pool.submit([this, b, nw] {
          std::vector<std::string> local(nw);
          for (auto& pr : *b) process_seq(pr.second, pr.first, local);
          for (size_t k = 0; k < nw; k++) flush(local, (int)k);
        });
        batch = std::make_shared<std::vector<std::pair<long, std::string>>>();
        batch_bytes = 0;
      };
      while (rd.next(line)) {
        size_t t = line.find('\t');  // toks[0] :448-449
        if (t != std::string::npos) line.resize(t);
        batch_bytes += line.size();
        batch->emplace_back(i, line);
        i++;
        if (batch_bytes > (1u << 18)) submit();
      }
      if (!batch->empty()) submit();
      pool.finish();
    }
    for (auto* f : outs) fclose(f);
  }
};

// sortBloom (cmd/muscato/main.go:318-385)
void sort_bloom(const Config& c) {
  for (size_t k = 0; k < c.Windows.size(); k++)
    sh("sort " + sort_args(c) + " -k1 " + tmp(c, "bmatch_" + std::to_string(k) + ".txt") + " > " +
       tmp(c, "smatch_" + std::to_string(k) + ".txt"));
}

// ---------------------------------------------------------------------------
// muscato_confirm (cmd/muscato_confirm/main.go)
// ---------------------------------------------------------------------------
struct Rec {
  std::string buf;
  std::vector<std::pair<uint32_t, uint32_t>> f;  // (offset, len) of tab-split fields (setfields :61-63)
  void setfields() {
    f.clear();
    uint32_t i = 0;
    while (true) {
      size_t j = buf.find('\t', i);
      if (j == std::string::npos) { f.emplace_back(i, (uint32_t)buf.size() - i); break; }
      f.emplace_back(i, (uint32_t)j - i);
      i = (uint32_t)j + 1;
    }
  }
  const char* p(int k) const { return buf.data() + f[k].first; }
  int n(int k) const { return (int)f[k].second; }
  int cmp0(const Rec& o) const {  // bytes.Compare on field 0
    int m = std::min(n(0), o.n(0));
    int r = memcmp(p(0), o.p(0), m);
    if (r) return r;
    return n(0) - o.n(0);
  }
};

// breader :69-148
struct BReader {
  LineReader rd;
  std::vector<Rec> recs;
  bool have_stash = false, done = false, any = false;
  Rec stash, last;
  explicit BReader(const std::string& path) : rd(path) {}
  bool next() {
    if (done) return false;
    recs.clear();
    if (have_stash) { recs.push_back(std::move(stash)); have_stash = false; }
    std::string line;
    for (int ii = 0; rd.next(line); ii++) {
      Rec rx;
      rx.buf = line;
      rx.setfields();
      any = true;
      if (!recs.empty() && recs[0].cmp0(rx) != 0) { stash = std::move(rx); have_stash = true; return true; }
      if (ii > 0 && last.cmp0(rx) > 0) die("file is not sorted");  // :130-135
      last = rx;
      recs.push_back(std::move(rx));
    }
    done = true;
    return true;
  }
};

struct QRect { int mismatch; std::string gob; };

// qinsert :424-448
void qinsert(std::vector<QRect>& q, QRect&& a, int max_matches) {
  q.push_back(std::move(a));
  size_t ii = q.size() - 1;
  while (ii > 0) {
    size_t jj = (ii - 1) / 2;
    if (q[jj].mismatch > q[ii].mismatch) { std::swap(q[jj], q[ii]); ii = jj; }
    else break;
  }
  if ((int)q.size() > max_matches) q.resize(max_matches);
}

inline int cdiff(const char* x, const char* y, int n) {  // :151-159
  int c = 0;
  for (int i = 0; i < n; i++) c += (x[i] != y[i]);
  return c;
}

struct Confirm {
  const Config& c;
  FILE* out;
  std::mutex mu;
  std::atomic<uint64_t> pairs{0}, kept{0};
  Confirm(const Config& cfg, FILE* o) : c(cfg), out(o) {}

  // searchpairs :171-250
  void searchpairs(const std::vector<Rec>& source, const std::vector<Rec>& match) {
    std::vector<QRect> qvals;
    bool first = c.MatchMode == "first";
    uint64_t np = 0;
    bool stop = false;
    for (const Rec& m : match) {
      if (stop) break;
      if (m.f.size() < 5) die("confirm: malformed match record");
      for (const Rec& s : source) {
        np++;
        int ltag = s.n(0), llft = s.n(1), lrgt = s.n(2);
        int nmiss = (int)((1 - c.PMatch) * (double)(ltag + llft + lrgt));  // :198 (float64, truncation)
        if (lrgt > m.n(2)) continue;                                         // :201-203
        int nx = cdiff(m.p(1), s.p(1), m.n(1));                              // :207 (ranges over mlft)
        nx += cdiff(m.p(2), s.p(2), lrgt);                                   // :208
        if (nx > nmiss) continue;
        long mposi = atol(std::string(m.p(4), m.n(4)).c_str());              // :214
        std::string g;
        g.reserve(2 * (ltag + llft + lrgt) + 40);
        g.append(s.p(1), llft).append(s.p(0), ltag).append(s.p(2), lrgt);
        g += '\t';
        g.append(m.p(1), m.n(1)).append(m.p(0), m.n(0)).append(m.p(2), lrgt);
        char t[64];
        int n = snprintf(t, sizeof t, "\t%ld\t%d\t", mposi - m.n(1), nx);
        g.append(t, n).append(m.p(3), m.n(3));
        g += '\n';
        if (first) {  // :233-238
          qvals.push_back(QRect{nx, std::move(g)});
          if ((int)qvals.size() > c.MaxMatches) { stop = true; break; }
        } else {
          qinsert(qvals, QRect{nx, std::move(g)}, c.MaxMatches);  // :241
        }
      }
    }
    pairs += np;
    kept += qvals.size();
    std::lock_guard<std::mutex> lk(mu);
    for (auto& v : qvals) fwrite(v.gob.data(), 1, v.gob.size(), out);
  }
};

void confirm_window(const Config& c, int win, int threads, uint64_t* pairs, uint64_t* kept) {
  std::string outname = tmp(c, "rmatch_" + std::to_string(win) + ".txt");
  FILE* out = fopen(outname.c_str(), "wb");
  if (!out) die("cannot create " + outname);
  BReader source(tmp(c, "win_" + std::to_string(win) + "_sorted.txt"));
  BReader match(tmp(c, "smatch_" + std::to_string(win) + ".txt"));
  Confirm cf(c, out);
  bool ms = source.next(), mb = match.next();
  if ((ms || mb) && !source.recs.empty() && !match.recs.empty()) {  // Q8: reference panics on empty input
    Pool pool(threads);
    while (true) {  // merge loop :375-416
      int cmp = source.recs[0].cmp0(match.recs[0]);
      if (cmp == 0) {
        auto s = std::make_shared<std::vector<Rec>>(source.recs);  // rcpy :262-271
        auto m = std::make_shared<std::vector<Rec>>(match.recs);
        pool.submit([&cf, s, m] { cf.searchpairs(*s, *m); });
        ms = source.next();
        mb = match.next();
        if (!(ms || mb)) break;
        if (!ms || !mb) {
          // One stream is exhausted: its recs still hold the (already joined) final block; keys of the
          // other stream only increase, so no further equal keys can occur.
          break;
        }
      } else if (cmp < 0) {
        ms = source.next();
        if (!ms) break;
      } else {
        mb = match.next();
        if (!mb) break;
      }
    }
    pool.finish();
  }
  fclose(out);
  *pairs += cf.pairs;
  *kept += cf.kept;
}

// confirm (cmd/muscato/main.go:387-420): groups of MaxConfirmProcs windows run concurrently.
void confirm_all(const Config& c, uint64_t* pairs, uint64_t* kept) {
  int nw = (int)c.Windows.size();
  for (int j = 0; j < nw;) {
    int m = std::min(nw, j + c.MaxConfirmProcs);
    int per = std::max(1, c.Threads / (m - j));
    std::vector<std::thread> th;
    std::vector<uint64_t> pp(m - j, 0), kk(m - j, 0);
    for (int k = j; k < m; k++) th.emplace_back([&, k] { confirm_window(c, k, per, &pp[k - j], &kk[k - j]); });
    for (auto& t : th) t.join();
    for (int k = j; k < m; k++) { *pairs += pp[k - j]; *kept += kk[k - j]; }
    j = m;
  }
}

// combineWindows (cmd/muscato/main.go:422-505): [combine_filter -> exact concat (Q9)] | sort -u | muscato_combine_windows
void combine_windows(const Config& c) {
  std::string files;
  for (size_t k = 0; k < c.Windows.size(); k++) files += " " + tmp(c, "rmatch_" + std::to_string(k) + ".txt");
  sh("cat" + files + " | sort " + sort_args(c) + " -u - > " + tmp(c, "rmatch_su.txt"));
  // muscato_combine_windows (cmd/muscato_combine_windows/main.go:94-143, writebest :36-60)
  LineReader rd(tmp(c, "rmatch_su.txt"));
  FILE* out = fopen(tmp(c, "matches.txt").c_str(), "wb");
  if (!out) die("cannot create matches.txt");
  std::vector<std::string> lines;
  std::vector<int> nm;
  std::string current, line;
  auto writebest = [&]() {
    int best = -1;
    for (int y : nm)
      if (best == -1 || y < best) best = y;
    for (size_t i = 0; i < lines.size(); i++)
      if (nm[i] <= best + c.MMTol) { fwrite(lines[i].data(), 1, lines[i].size(), out); fputc('\n', out); }
  };
  while (rd.next(line)) {
    auto f = fields(line);
    if (f.size() < 4) die("combine_windows: malformed line");
    if (current.empty() || f[0] == current) {
      lines.push_back(line);
      nm.push_back(atoi(f[3].c_str()));
      current = f[0];
      continue;
    }
    writebest();
    lines.clear();
    nm.clear();
    lines.push_back(line);
    nm.push_back(atoi(f[3].c_str()));
    current = f[0];
  }
  writebest();
  fclose(out);
}

// sortByGeneId + joinGeneNames + joinReadNames (cmd/muscato/main.go:507-676)
void epilogue(const Config& c) {
  sh("sort " + sort_args(c) + " -k5 " + tmp(c, "matches.txt") + " > " + tmp(c, "matches_sg.txt"));
  sh("join -1 5 -2 1 -t $'\\t' " + tmp(c, "matches_sg.txt") + " " + c.GeneIdFileName +
     " | cut -d $'\\t' -f1 --complement - > " + tmp(c, "matches_sn.txt"));
  sh("sort -k1 " + sort_args(c) + " " + tmp(c, "matches_sn.txt") + " > " + tmp(c, "matches_sn_sorted.txt"));
  sh("join -1 1 -2 1 -t $'\\t' " + tmp(c, "matches_sn_sorted.txt") + " " + tmp(c, "reads_sorted.txt") + " > " +
     c.ResultsFileName);
}

// muscato_nonmatch (cmd/muscato_nonmatch/main.go:40-114) with an exact set (Q10)
void nonmatch(const Config& c) {
  std::unordered_set<std::string> seen;
  {
    LineReader rd(c.ResultsFileName);
    std::string line;
    while (rd.next(line)) {
      auto f = fields(line);
      if (!f.empty()) seen.insert(f[0]);
    }
  }
  // Output name :66-71
  std::string a, b = c.ResultsFileName;
  size_t sl = b.rfind('/');
  if (sl != std::string::npos) { a = b.substr(0, sl + 1); b = b.substr(sl + 1); }
  std::vector<std::string> parts;
  {
    size_t i = 0;
    while (true) {
      size_t j = b.find('.', i);
      if (j == std::string::npos) { parts.emplace_back(b, i); break; }
      parts.emplace_back(b, i, j - i);
      i = j + 1;
    }
  }
  std::string d = parts.back();
  parts.back() = "nonmatch";
  parts.push_back(d + ".fastq");
  std::string outname = a;
  for (size_t i = 0; i < parts.size(); i++) { if (i) outname += '.'; outname += parts[i]; }
  FILE* out = fopen(outname.c_str(), "wb");
  if (!out) die("cannot create " + outname);
  LineReader rd(tmp(c, "reads_sorted.txt"));
  std::string line;
  while (rd.next(line)) {
    auto f = fields(line);
    if (f.size() < 3) die("nonmatch: malformed reads_sorted line");
    if (seen.count(f[0])) continue;
    fprintf(out, "%s#%s\n%s\n+\n%s\n", f[2].c_str(), f[1].c_str(), f[0].c_str(), std::string(f[0].size(), '!').c_str());
  }
  fclose(out);
}

void upstream(const Config& c) {
  prep_reads(c);
  window_reads(c);
  sort_windows(c);
}

void hotpath(const Config& c, bool report) {
  double t0 = now_s();
  Screen scr(c);
  scr.build_bloom();
  double t1 = now_s();
  scr.search();
  double t2 = now_s();
  sort_bloom(c);
  double t3 = now_s();
  uint64_t pairs = 0, kept = 0;
  confirm_all(c, &pairs, &kept);
  double t4 = now_s();
  combine_windows(c);
  double t5 = now_s();
  if (report)
    printf("{\"threads\": %d, \"target_bases\": %llu, \"candidates\": %llu, \"pairs\": %llu, \"kept\": %llu, "
           "\"bloom_build_s\": %.6f, \"screen_s\": %.6f, \"sort_s\": %.6f, \"confirm_s\": %.6f, \"combine_s\": %.6f, "
           "\"total_s\": %.6f}\n",
           c.Threads, (unsigned long long)scr.bases.load(), (unsigned long long)scr.recs.load(),
           (unsigned long long)pairs, (unsigned long long)kept, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0);
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) die("usage: muscato_oracle <prep_targets|pipeline|upstream|hotpath|epilogue> ...");
  std::string cmd = argv[1];
  if (cmd == "prep_targets") {
    bool rev = false;
    int a = 2;
    if (a < argc && std::string(argv[a]) == "-rev") { rev = true; a++; }
    if (argc - a != 3) die("usage: prep_targets [-rev] <in> <seq_out> <ids_out>");
    std::string in = argv[a], lower = in;
    for (auto& ch : lower) ch = (char)tolower(ch);
    bool fasta = lower.size() >= 5 && lower.compare(lower.size() - 5, 5, "fasta") == 0;  // :321-322
    prep_targets(in, argv[a + 1], argv[a + 2], rev, fasta);
    return 0;
  }
  if (argc != 3) die("usage: muscato_oracle " + cmd + " <config.json>");
  Config c = read_config(argv[2]);
  if (c.TempDir.empty()) die("TempDir must be set");
  if (cmd == "pipeline") {
    upstream(c);
    hotpath(c, false);
    epilogue(c);
    nonmatch(c);
  } else if (cmd == "upstream") {
    upstream(c);
  } else if (cmd == "prep_reads") {  // prepReads alone: TempDir/reads_sorted.txt
    prep_reads(c);
  } else if (cmd == "windows") {  // windowReads + sortWindows on an existing TempDir/reads_sorted.txt
    window_reads(c);
    sort_windows(c);
  } else if (cmd == "hotpath") {
    hotpath(c, true);
  } else if (cmd == "epilogue") {
    epilogue(c);
    nonmatch(c);
  } else {
    die("unknown sub-command " + cmd);
  }
  return 0;
}
