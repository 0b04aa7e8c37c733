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

namespace {

// One entry header at buf[0 .. avail): key, binary float matrix marker, sizes.  `remaining` = bytes from buf[0] to the end of the
// archive (>= avail: the payload need not be in the buffer).  Returns 1 and fills the outputs (offsets relative to buf) when a
// whole 'FM ' / 'DM ' entry starts here and fits; 0 when what starts here is something else (text, compressed, malformed,
// truncated); -1 when `avail` is too short to tell (the caller reads more).
inline int ark_parse_header(const uint8_t* buf, int64_t avail, int64_t remaining, int64_t* key_off, int32_t* key_len, int32_t* rows,
                            int32_t* cols, int32_t* elem_bytes, int64_t* payload_off) {
  const void* sp = memchr(buf, ' ', size_t(avail));
  if (!sp) return avail < remaining && avail <= 4200 ? -1 : 0;
  int64_t k0 = 0, k1 = static_cast<const uint8_t*>(sp) - buf;
  while (k0 < k1 && ark_space(buf[k0])) ++k0;
  while (k1 > k0 && ark_space(buf[k1 - 1])) --k1;
  if (k1 == k0 || k1 - k0 > 4096) return 0;
  for (int64_t i = k0; i < k1; ++i)
    if (!ark_key_char(buf[i])) return 0;
  const int64_t h = (static_cast<const uint8_t*>(sp) - buf) + 1;          // first byte after the separating space
  if (h + 15 > remaining) return 0;
  if (h + 15 > avail) return -1;
  if (buf[h] != 0 || buf[h + 1] != 'B' || (buf[h + 2] != 'F' && buf[h + 2] != 'D') || buf[h + 3] != 'M' || buf[h + 4] != ' ' ||
      buf[h + 5] != 4 || buf[h + 10] != 4)
    return 0;
  int32_t r, c;
  memcpy(&r, buf + h + 6, 4);
  memcpy(&c, buf + h + 11, 4);
  const int32_t eb = buf[h + 2] == 'F' ? 4 : 8;
  if (r < 0 || c < 0) return 0;
  const int64_t payload = h + 15, bytes = int64_t(r) * c * eb;
  if (payload + bytes > remaining) return 0;
  *key_off = k0; *key_len = int32_t(k1 - k0);
  *rows = r; *cols = c; *elem_bytes = eb; *payload_off = payload;
  return 1;
}

}  // namespace

extern "C" int64_t xv_ark_scan(const uint8_t* buf, int64_t len, int64_t max_entries, int64_t* key_off, int32_t* key_len,
                               int32_t* rows, int32_t* cols, int32_t* elem_bytes, int64_t* payload_off, int64_t* consumed) {
  int64_t pos = 0, n = 0;
  if (!buf || len < 0 || !key_off || !key_len || !rows || !cols || !elem_bytes || !payload_off || !consumed) return -1;
  while (n < max_entries && pos < len) {
    int64_t ko, po;
    if (ark_parse_header(buf + pos, len - pos, len - pos, &ko, key_len + n, rows + n, cols + n, elem_bytes + n, &po) != 1) break;
    key_off[n] = pos + ko;
    payload_off[n] = pos + po;
    pos = payload_off[n] + int64_t(rows[n]) * cols[n] * elem_bytes[n];
    ++n;
  }
  *consumed = pos;
  return n;
}
