// szio.hpp -- Snappy *framing format* reader / writer for the reference's ".sz" files
// (golang/snappy NewReader / NewBufferedWriter, used by every stage: e.g.
// cmd/muscato_screen/main.go:126, :378).  Stream id "\xff\x06\x00\x00sNaPpY", then chunks
// type(1) len(3 LE) [masked CRC32C(4) payload]; 0x00 = Snappy block, 0x01 = stored.
// The writer emits stored chunks, which every conforming reader (golang/snappy, sztool)
// accepts.  Container format only: the content is newline-delimited text.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
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

inline uint32_t crc32c(const uint8_t* p, size_t n) {
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

inline std::string decompress(const std::string& raw) {
  std::string out;
  const uint8_t* p = reinterpret_cast<const uint8_t*>(raw.data());
  size_t i = 0, n = raw.size();
  bool magic = false;
  while (i < n) {
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
      const uint32_t want = body[0] | ((uint32_t)body[1] << 8) | ((uint32_t)body[2] << 16) | ((uint32_t)body[3] << 24);
      const size_t before = out.size();
      if (type == 0x00) snappy_block_decode(body + 4, len - 4, out);
      else out.append(reinterpret_cast<const char*>(body + 4), len - 4);
      if (masked_crc(reinterpret_cast<const uint8_t*>(out.data()) + before, out.size() - before) != want)
        throw std::runtime_error("sz: CRC mismatch");
    } else if (type >= 0x02 && type <= 0x7F) {
      throw std::runtime_error("sz: reserved unskippable chunk");
    }  // 0x80..0xfe: skippable / padding
  }
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
