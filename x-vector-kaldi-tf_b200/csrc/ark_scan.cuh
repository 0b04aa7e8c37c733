// ark_scan.cuh -- host-only: index of the binary matrix entries of a Kaldi ark that lies in memory (an mmap'ed file).
// The extractor's reader walks this index instead of parsing one header at a time in Python, and fetches the payloads
// with pread from several threads (models.py: _read_batches).  Layout parsed (reference kaldi_io.py:120-133, :413-437):
//     <key> ' '  '\0' 'B'  'F'|'D' 'M' ' '  '\4' <int32 rows>  '\4' <int32 cols>  rows * cols * (4|8) bytes
// Anything else (text matrices, compressed 'CM' matrices, a malformed key, a truncated payload) ends the scan: the caller
// continues from *consumed with the general Python parser, which also owns the error messages.
#pragma once
#include <stdint.h>
#include <string.h>

#include "../../include/xvec.h"

namespace {

inline bool ark_key_char(uint8_t c) {
  return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '.' || c == '/' || c == '_' || c == '-';
}
inline bool ark_space(uint8_t c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

}  // namespace

extern "C" int64_t xv_ark_scan(const uint8_t* buf, int64_t len, int64_t max_entries, int64_t* key_off, int32_t* key_len,
                               int32_t* rows, int32_t* cols, int32_t* elem_bytes, int64_t* payload_off, int64_t* consumed) {
  int64_t pos = 0, n = 0;
  if (!buf || len < 0 || !key_off || !key_len || !rows || !cols || !elem_bytes || !payload_off || !consumed) return -1;
  while (n < max_entries && pos < len) {
    // key: bytes up to the first ' ', white space stripped from both ends
    const void* sp = memchr(buf + pos, ' ', size_t(len - pos));
    if (!sp) break;
    int64_t k0 = pos, k1 = static_cast<const uint8_t*>(sp) - buf;
    while (k0 < k1 && ark_space(buf[k0])) ++k0;
    while (k1 > k0 && ark_space(buf[k1 - 1])) --k1;
    if (k1 == k0 || k1 - k0 > 4096) break;
    bool ok = true;
    for (int64_t i = k0; i < k1 && ok; ++i) ok = ark_key_char(buf[i]);
    if (!ok) break;
    int64_t h = (static_cast<const uint8_t*>(sp) - buf) + 1;          // first byte after the separating space
    if (h + 15 > len) break;
    if (buf[h] != 0 || buf[h + 1] != 'B' || (buf[h + 2] != 'F' && buf[h + 2] != 'D') || buf[h + 3] != 'M' || buf[h + 4] != ' ' ||
        buf[h + 5] != 4 || buf[h + 10] != 4)
      break;
    int32_t r, c;
    memcpy(&r, buf + h + 6, 4);
    memcpy(&c, buf + h + 11, 4);
    const int32_t eb = buf[h + 2] == 'F' ? 4 : 8;
    if (r < 0 || c < 0) break;
    const int64_t payload = h + 15, bytes = int64_t(r) * c * eb;
    if (payload + bytes > len) break;
    key_off[n] = k0; key_len[n] = int32_t(k1 - k0);
    rows[n] = r; cols[n] = c; elem_bytes[n] = eb; payload_off[n] = payload;
    ++n;
    pos = payload + bytes;
  }
  *consumed = pos;
  return n;
}
