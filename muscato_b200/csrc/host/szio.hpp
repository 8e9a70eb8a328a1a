// szio.hpp -- Snappy *framing format* reader / writer for the reference's ".sz" files
// (golang/snappy NewReader / NewBufferedWriter, used by every stage: e.g.
// cmd/muscato_screen/main.go:126, :378).  Stream id "\xff\x06\x00\x00sNaPpY", then chunks
// type(1) len(3 LE) [masked CRC32C(4) payload]; 0x00 = Snappy block, 0x01 = stored.
// The writer emits stored chunks, which every conforming reader (golang/snappy, sztool)
// accepts.  Container format only: the content is newline-delimited text.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace szio {

inline const uint32_t* crc_table() {
  static uint32_t tab[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; i++) {
      uint32_t c = i;
      for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;  // CRC-32C, reflected
      tab[i] = c;
    }
    init = true;
  }
  return tab;
}

#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("sse4.2"))) inline uint32_t crc32c_hw(const uint8_t* p, size_t n) {
  uint64_t c = 0xFFFFFFFFu;
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    uint64_t v;
    memcpy(&v, p + i, 8);
    c = __builtin_ia32_crc32di(c, v);
  }
  uint32_t c32 = (uint32_t)c;
  for (; i < n; i++) c32 = __builtin_ia32_crc32qi(c32, p[i]);
  return c32 ^ 0xFFFFFFFFu;
}
#endif

inline uint32_t crc32c(const uint8_t* p, size_t n) {
#if defined(__x86_64__) && defined(__GNUC__)
  static const bool hw = __builtin_cpu_supports("sse4.2");
  if (hw) return crc32c_hw(p, n);  // the CRC32 instruction computes exactly CRC-32C
#endif
  const uint32_t* tab = crc_table();
  uint32_t c = 0xFFFFFFFFu;
  for (size_t i = 0; i < n; i++) c = tab[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}

inline uint32_t masked_crc(const uint8_t* p, size_t n) {
  const uint32_t c = crc32c(p, n);
  return ((c >> 15) | (c << 17)) + 0xA282EAD8u;
}

// Raw Snappy block: varint uncompressed length, then literal / copy elements.
inline void snappy_block_decode(const uint8_t* b, size_t n, std::string& out) {
  size_t i = 0;
  uint64_t ulen = 0;
  int shift = 0;
  while (true) {
    if (i >= n) throw std::runtime_error("snappy: truncated length");
    const uint8_t c = b[i++];
    ulen |= (uint64_t)(c & 0x7F) << shift;
    if (c < 0x80) break;
    shift += 7;
  }
  const size_t start = out.size();
  out.reserve(start + ulen);
  while (i < n) {
    const uint8_t tag = b[i++];
    const int kind = tag & 3;
    if (kind == 0) {
      size_t len = tag >> 2;
      if (len >= 60) {
        const int nb = (int)len - 59;
        if (i + nb > n) throw std::runtime_error("snappy: truncated literal length");
        len = 0;
        for (int k = 0; k < nb; k++) len |= (size_t)b[i + k] << (8 * k);
        i += nb;
      }
      len += 1;
      if (i + len > n) throw std::runtime_error("snappy: truncated literal");
      out.append(reinterpret_cast<const char*>(b + i), len);
      i += len;
      continue;
    }
    size_t len, off;
    if (kind == 1) {
      if (i + 1 > n) throw std::runtime_error("snappy: truncated copy");
      len = 4 + ((tag >> 2) & 7);
      off = ((size_t)(tag >> 5) << 8) | b[i];
      i += 1;
    } else if (kind == 2) {
      if (i + 2 > n) throw std::runtime_error("snappy: truncated copy");
      len = 1 + (tag >> 2);
      off = b[i] | ((size_t)b[i + 1] << 8);
      i += 2;
    } else {
      if (i + 4 > n) throw std::runtime_error("snappy: truncated copy");
      len = 1 + (tag >> 2);
      off = b[i] | ((size_t)b[i + 1] << 8) | ((size_t)b[i + 2] << 16) | ((size_t)b[i + 3] << 24);
      i += 4;
    }
    if (off == 0 || off > out.size() - start) throw std::runtime_error("snappy: bad copy offset");
    const size_t from = out.size() - off;
    for (size_t k = 0; k < len; k++) out.push_back(out[from + k]);  // may overlap: byte by byte
  }
  if (out.size() - start != ulen) throw std::runtime_error("snappy: length mismatch");
}

inline std::string read_all(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot open " + path);
  std::string raw;
  char buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) raw.append(buf, n);
  fclose(f);
  return raw;
}

inline bool is_framed(const std::string& raw) {
  return raw.size() >= 10 && memcmp(raw.data(), "\xff\x06\x00\x00sNaPpY", 10) == 0;
}

// Raw Snappy block decoded into a buffer of exactly `ulen` bytes (the length its header states).
inline void snappy_block_decode_to(const uint8_t* b, size_t n, uint8_t* dst, size_t ulen) {
  size_t i = 0;
  while (i < n && (b[i] & 0x80)) i++;  // skip the varint length (validated by the caller)
  i++;
  size_t o = 0;
  while (i < n) {
    const uint8_t tag = b[i++];
    const int kind = tag & 3;
    if (kind == 0) {
      size_t len = tag >> 2;
      if (len >= 60) {
        const int nb = (int)len - 59;
        if (i + nb > n) throw std::runtime_error("snappy: truncated literal length");
        len = 0;
        for (int k = 0; k < nb; k++) len |= (size_t)b[i + k] << (8 * k);
        i += nb;
      }
      len += 1;
      if (i + len > n || o + len > ulen) throw std::runtime_error("snappy: truncated literal");
      memcpy(dst + o, b + i, len);
      o += len;
      i += len;
      continue;
    }
    size_t len, off;
    if (kind == 1) {
      if (i + 1 > n) throw std::runtime_error("snappy: truncated copy");
      len = 4 + ((tag >> 2) & 7);
      off = ((size_t)(tag >> 5) << 8) | b[i];
      i += 1;
    } else if (kind == 2) {
      if (i + 2 > n) throw std::runtime_error("snappy: truncated copy");
      len = 1 + (tag >> 2);
      off = b[i] | ((size_t)b[i + 1] << 8);
      i += 2;
    } else {
      if (i + 4 > n) throw std::runtime_error("snappy: truncated copy");
      len = 1 + (tag >> 2);
      off = b[i] | ((size_t)b[i + 1] << 8) | ((size_t)b[i + 2] << 16) | ((size_t)b[i + 3] << 24);
      i += 4;
    }
    if (off == 0 || off > o || o + len > ulen) throw std::runtime_error("snappy: bad copy");
    if (off >= len) memcpy(dst + o, dst + o - off, len);
    else for (size_t k = 0; k < len; k++) dst[o + k] = dst[o - off + k];  // overlapping run: byte by byte
    o += len;
  }
  if (o != ulen) throw std::runtime_error("snappy: length mismatch");
}

// Un-frame a ".sz" stream.  threads = 0: all host threads.
inline std::string decompress(const std::string& raw, unsigned threads = 0) {
  struct Chunk {
    uint8_t type;
    const uint8_t* body;  // after the 4 CRC bytes
    size_t len, off, ulen;
    uint32_t want;
  };
  std::vector<Chunk> chunks;
  const uint8_t* p = reinterpret_cast<const uint8_t*>(raw.data());
  size_t i = 0, n = raw.size(), total = 0;
  bool magic = false;
  while (i < n) {  // pass 1: index the chunks and their uncompressed lengths
    if (i + 4 > n) throw std::runtime_error("sz: truncated chunk header");
    const uint8_t type = p[i];
    const size_t len = p[i + 1] | ((size_t)p[i + 2] << 8) | ((size_t)p[i + 3] << 16);
    if (i + 4 + len > n) throw std::runtime_error("sz: truncated chunk");
    const uint8_t* body = p + i + 4;
    i += 4 + len;
    if (type == 0xFF) {
      if (len != 6 || memcmp(body, "sNaPpY", 6) != 0) throw std::runtime_error("sz: bad stream identifier");
      magic = true;
      continue;
    }
    if (!magic) throw std::runtime_error("sz: missing stream identifier");
    if (type == 0x00 || type == 0x01) {
      if (len < 4) throw std::runtime_error("sz: short chunk");
      Chunk c;
      c.type = type;
      c.body = body + 4;
      c.len = len - 4;
      c.want = body[0] | ((uint32_t)body[1] << 8) | ((uint32_t)body[2] << 16) | ((uint32_t)body[3] << 24);
      c.off = total;
      c.ulen = c.len;
      if (type == 0x00) {
        uint64_t ulen = 0;
        int shift = 0;
        size_t k = 0;
        while (true) {
          if (k >= c.len || shift > 35) throw std::runtime_error("snappy: truncated length");
          const uint8_t b = c.body[k++];
          ulen |= (uint64_t)(b & 0x7F) << shift;
          if (b < 0x80) break;
          shift += 7;
        }
        c.ulen = (size_t)ulen;
      }
      total += c.ulen;
      chunks.push_back(c);
    } else if (type >= 0x02 && type <= 0x7F) {
      throw std::runtime_error("sz: reserved unskippable chunk");
    }  // 0x80..0xfe: skippable / padding
  }
  std::string out(total, '\0');
  uint8_t* o = reinterpret_cast<uint8_t*>(&out[0]);
  // pass 2: decode + checksum, chunks dealt to the threads through one atomic counter
  std::atomic<size_t> next{0};
  std::atomic<bool> failed{false};
  std::string err;
  auto work = [&]() {
    try {
      while (!failed.load(std::memory_order_relaxed)) {
        const size_t k = next.fetch_add(16);
        if (k >= chunks.size()) break;
        for (size_t c = k; c < std::min(chunks.size(), k + 16); c++) {
          const Chunk& ch = chunks[c];
          if (ch.type == 0x00) snappy_block_decode_to(ch.body, ch.len, o + ch.off, ch.ulen);
          else memcpy(o + ch.off, ch.body, ch.len);
          if (masked_crc(o + ch.off, ch.ulen) != ch.want) throw std::runtime_error("sz: CRC mismatch");
        }
      }
    } catch (const std::exception& e) {
      if (!failed.exchange(true)) err = e.what();
    }
  };
  unsigned nt = threads ? threads : std::max(1u, std::thread::hardware_concurrency());
  nt = (unsigned)std::min<size_t>(nt, (chunks.size() + 63) / 64 + 1);  // small files: no thread start-up cost
  if (nt <= 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; t++) pool.emplace_back(work);
    for (auto& t : pool) t.join();
  }
  if (failed.load()) throw std::runtime_error(err);
  return out;
}

// Text of a file, transparently un-framing ".sz".
inline std::string read_text(const std::string& path) {
  std::string raw = read_all(path);
  return is_framed(raw) ? decompress(raw) : raw;
}

inline void write_file(const std::string& path, const std::string& data, bool framed) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) throw std::runtime_error("cannot create " + path);
  if (!framed) {
    fwrite(data.data(), 1, data.size(), f);
  } else {
    fwrite("\xff\x06\x00\x00sNaPpY", 1, 10, f);
    const size_t kBlock = 65536;
    for (size_t i = 0; i < data.size(); i += kBlock) {
      const size_t n = std::min(kBlock, data.size() - i);
      const uint8_t* p = reinterpret_cast<const uint8_t*>(data.data()) + i;
      const uint32_t crc = masked_crc(p, n);
      const uint32_t len = (uint32_t)n + 4;
      const uint8_t hdr[8] = {0x01, (uint8_t)len, (uint8_t)(len >> 8), (uint8_t)(len >> 16),
                              (uint8_t)crc, (uint8_t)(crc >> 8), (uint8_t)(crc >> 16), (uint8_t)(crc >> 24)};
      fwrite(hdr, 1, 8, f);
      fwrite(p, 1, n, f);
    }
  }
  if (fclose(f) != 0) throw std::runtime_error("write failed: " + path);
}

inline bool ends_with(const std::string& s, const char* suf) {
  const size_t n = strlen(suf);
  return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

}  // namespace szio
