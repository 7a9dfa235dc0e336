// FASTQ record-boundary scan for the batcher: the part of seq_io 0.3.2's reader that
// FastqReader::read_record_set (fastq.rs:166-172) relies on.  The fields themselves (head / seq /
// qual, CR stripping) are cut on the device; the host only needs where records start, how long
// the longest one is, and whether the text is well-formed.
#include <string.h>

#include "../../include/barkit_b200.h"

extern "C" int bk_tokenize(const uint8_t* text, uint64_t len, int final_chunk, uint32_t* rec_off,
                           uint32_t cap_records, uint32_t* n_out, uint64_t* consumed,
                           uint32_t* max_rec_bytes) {
  if (!text || !rec_off || !n_out) return BK_E_ARG;
  uint64_t pos = 0;
  uint32_t n = 0, maxrec = 0;
  int rc = BK_OK;
  while (pos < len && n < cap_records) {
    if (pos >= (1ull << 32) - 1) break;  // offsets are 32-bit; caller continues with a new batch
    if (text[pos] == '\n' || text[pos] == '\r') {  // blank lines after the last record
      uint64_t q = pos;
      while (q < len && (text[q] == '\n' || text[q] == '\r')) ++q;
      if (q == len && final_chunk) { pos = len; break; }
      rc = BK_E_FORMAT;
      break;
    }
    uint64_t p = pos;
    uint64_t line_b[4], line_e[4];
    bool complete = true;
    for (int l = 0; l < 4; ++l) {
      line_b[l] = p;
      if (p > len) { complete = false; break; }
      const void* nl = p < len ? memchr(text + p, '\n', len - p) : nullptr;
      if (!nl) {
        // last line without EOL: only acceptable for the quality line of the final record
        if (!(final_chunk && l == 3 && p <= len)) { complete = false; break; }
        line_e[l] = len;
        p = len;
      } else {
        line_e[l] = (const uint8_t*)nl - text;
        p = line_e[l] + 1;
      }
    }
    if (!complete) {
      if (final_chunk) rc = BK_E_FORMAT;  // truncated record at end of input
      break;
    }
    if (p > (1ull << 32) - 1) break;
    if (text[line_b[0]] != '@' || line_b[2] >= len || text[line_b[2]] != '+') { rc = BK_E_FORMAT; break; }
    uint64_t sl = line_e[1] - line_b[1], ql = line_e[3] - line_b[3];
    if (sl && text[line_e[1] - 1] == '\r') --sl;
    if (ql && text[line_e[3] - 1] == '\r') --ql;
    if (sl != ql) { rc = BK_E_FORMAT; break; }
    rec_off[n++] = (uint32_t)pos;
    uint64_t rb = p - pos;
    if (rb > maxrec) maxrec = (uint32_t)rb;
    pos = p;
  }
  rec_off[n] = (uint32_t)pos;
  *n_out = n;
  if (consumed) *consumed = pos;
  if (max_rec_bytes) *max_rec_bytes = maxrec;
  return rc;
}
