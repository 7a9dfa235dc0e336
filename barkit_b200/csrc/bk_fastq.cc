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

// ---- a mate without a pattern never has to cross the bus ------------------------------------------
//
// In paired mode a mate whose pattern is `None` (run.rs:233-245) is never rewritten: get_new_reads
// (run.rs:149-160) copies its record when the other mate matched and drops it otherwise, and the
// writer (fastq.rs:259-263) frames it as `@head LF seq LF + LF qual LF`.  The host already holds
// that text, so bk_submit sends only the other mate to the device and bk_wait assembles this mate's
// output here from the caller's input buffer and the device's match table: kept records in input
// order, runs of consecutive kept records that already have the writer's framing as one memcpy.
#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace bk {

namespace {

inline const uint8_t* find_lf(const uint8_t* p, const uint8_t* e) {
  const void* q = p < e ? memchr(p, '\n', (size_t)(e - p)) : nullptr;
  return q ? (const uint8_t*)q : e;
}
inline const uint8_t* strip_cr(const uint8_t* b, const uint8_t* e) { return (e > b && e[-1] == '\r') ? e - 1 : e; }

// Copy of a long run with streaming stores: the output is written once and not read again here, so
// fetching its cache lines first (what plain stores do) would add a third of the pass's memory traffic.
inline void copy_stream(uint8_t* dst, const uint8_t* src, size_t n) {
#if defined(__SSE2__)
  if (n >= 256) {
    // (to a CACHE LINE boundary: a write-combining buffer that is handed its 64 bytes by four stores of one turn goes out
    // as one full-line write; at 16-byte alignment every turn straddles two lines and one core copies at half the rate --
    // 71 against 35 ms per 1 M records of 329 bytes, tools/gather_bench.py)
    const size_t head = (64 - ((uintptr_t)dst & 63)) & 63;
    memcpy(dst, src, head);
    dst += head; src += head; n -= head;
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
      const __m128i a = _mm_loadu_si128((const __m128i*)(src + i)), b = _mm_loadu_si128((const __m128i*)(src + i + 16));
      const __m128i c = _mm_loadu_si128((const __m128i*)(src + i + 32)), d = _mm_loadu_si128((const __m128i*)(src + i + 48));
      _mm_stream_si128((__m128i*)(dst + i), a);
      _mm_stream_si128((__m128i*)(dst + i + 16), b);
      _mm_stream_si128((__m128i*)(dst + i + 32), c);
      _mm_stream_si128((__m128i*)(dst + i + 48), d);
    }
    memcpy(dst + i, src + i, n - i);
    return;
  }
#endif
  memcpy(dst, src, n);
}

// the record [b, e) is byte for byte what the writer would produce (parse_record's closed form, bk_kernels.cuh)
inline bool writer_framed(const uint8_t* b, const uint8_t* e) {
  const uint8_t* nl1 = find_lf(b, e);
  if (nl1 == e || nl1 == b) return false;
  const uint8_t* sb = nl1 + 1;
  const size_t rem = (size_t)(e - sb);
  if (rem < 4 || (rem & 1)) return false;
  const size_t L = (rem - 4) >> 1;
  if (!(sb[L] == '\n' && sb[L + 1] == '+' && sb[L + 2] == '\n' && e[-1] == '\n')) return false;
  if (L != 0 && (sb[L - 1] == '\r' || e[-2] == '\r')) return false;
  return !(nl1 > b + 1 && nl1[-1] == '\r');
}

// The same answer from eight byte probes when the header is `H` bytes long, as the previous record's was (the headers
// of one file rarely differ in length): `@` H bytes LF | L bytes LF | + LF | L bytes LF with L = (e - b - H - 6) / 2.
// A record holds exactly four line ends (its bounds are every fourth one of the text, bk_tokenize), so four found at
// distinct places are all of them and the lines are as assumed.  False = "not shown", not "not framed": the caller
// falls back to writer_framed, which scans the header.
inline bool writer_framed_probe(const uint8_t* b, const uint8_t* e, size_t H) {
  const size_t len = (size_t)(e - b);
  if (H + 6 > len || ((len - H) & 1)) return false;
  const size_t L = (len - H - 6) >> 1;
  const uint8_t* sb = b + H + 2;
  if (!(sb[-1] == '\n' && sb[L] == '\n' && sb[L + 1] == '+' && sb[L + 2] == '\n' && e[-1] == '\n')) return false;
  if (L != 0 && (sb[L - 1] == '\r' || e[-2] == '\r')) return false;
  return !(H != 0 && sb[-2] == '\r');
}
constexpr uint32_t kAhead = 12;  // records between the prefetch and the probes (gather_kept)

// any other shape: cut the four lines as seq_io does (CR stripped, text after '+' dropped) and re-frame
inline uint8_t* reframe(const uint8_t* b, const uint8_t* e, uint8_t* out) {
  const uint8_t* nl1 = find_lf(b, e);
  const uint8_t* hb = b + 1 <= e ? b + 1 : e;
  const uint8_t* he = strip_cr(hb, nl1 > hb ? nl1 : hb);
  const uint8_t* sb = nl1 + 1 < e ? nl1 + 1 : e;
  const uint8_t* nl2 = find_lf(sb, e);
  const uint8_t* se = strip_cr(sb, nl2);
  const uint8_t* pb = nl2 + 1 < e ? nl2 + 1 : e;
  const uint8_t* nl3 = find_lf(pb, e);
  const uint8_t* qb = nl3 + 1 < e ? nl3 + 1 : e;
  const uint8_t* nl4 = find_lf(qb, e);
  const uint8_t* qe = strip_cr(qb, nl4);
  *out++ = '@';
  memcpy(out, hb, (size_t)(he - hb)); out += he - hb;
  *out++ = '\n';
  memcpy(out, sb, (size_t)(se - sb)); out += se - sb;
  *out++ = '\n'; *out++ = '+'; *out++ = '\n';
  memcpy(out, qb, (size_t)(qe - qb)); out += qe - qb;
  *out++ = '\n';
  return out;
}

}  // namespace

// Kept records (keep[i] >= 0) of text/off[0..n] -> out, in the writer's framing.  Returns the bytes
// written, or UINT64_MAX when `cap` cannot hold them.  A re-framed record is never longer than its
// input, except that the batch's very last record gains one byte when it had no EOL; so every thread
// can write at the offset its kept records' input lengths give, and if some record came out shorter
// the parts are closed up afterwards.
uint64_t gather_kept(const uint8_t* text, const uint32_t* off, uint32_t n, const int32_t* keep, uint8_t* out,
                     uint64_t cap, int max_threads) {
  if (n == 0) return 0;
  int T = n < 50000 ? 1 : std::max(1, std::min(max_threads, 32));
  std::vector<uint64_t> raw((size_t)T + 1, 0), wrote((size_t)T, 0), start((size_t)T, 0);
  std::vector<uint32_t> lo((size_t)T + 1);
  for (int t = 0; t <= T; ++t) lo[t] = (uint32_t)((uint64_t)n * t / T);
  auto count = [&](int t) {
    uint64_t s = 0;
    for (uint32_t i = lo[t]; i < lo[t + 1]; ++i)
      if (keep[i] >= 0) s += off[i + 1] - off[i];
    raw[t + 1] = s;
  };
  auto copy = [&](int t) {
    uint8_t* o = out + start[t];
    uint8_t* const o0 = o;
    uint32_t i = lo[t];
    const uint32_t hi = lo[t + 1];
    size_t hguess = 0;  // header length of the last record whose header was scanned
    while (i < hi) {
      if (keep[i] < 0) { ++i; continue; }
      // maximal run of kept, writer-framed records: one copy
      uint32_t j = i;
      while (j < hi && keep[j] >= 0) {
        const uint8_t* const b = text + off[j];
        const uint8_t* const e = text + off[j + 1];
#if defined(__SSE2__)
        if (j + kAhead < hi) {  // the probes touch a record at three places: have those lines on their way
          const uint8_t* const pb = text + off[j + kAhead];
          const size_t pl = off[j + kAhead + 1] - off[j + kAhead];
          _mm_prefetch((const char*)pb, _MM_HINT_T0);
          _mm_prefetch((const char*)pb + (pl >> 1), _MM_HINT_T0);
          _mm_prefetch((const char*)pb + pl - 1, _MM_HINT_T0);
        }
#endif
        if (!writer_framed_probe(b, e, hguess)) {
          if (!writer_framed(b, e)) break;
          hguess = (size_t)(find_lf(b, e) - b) - 1;  // (framed: the header line is there and ends with LF)
        }
        ++j;
      }
      if (j > i) {
        copy_stream(o, text + off[i], off[j] - off[i]);
        o += off[j] - off[i];
        i = j;
      } else {
        o = reframe(text + off[i], text + off[i + 1], o);
        ++i;
      }
    }
#if defined(__SSE2__)
    _mm_sfence();
#endif
    wrote[t] = (uint64_t)(o - o0);
  };
  // one set of threads for both passes: every thread counts its part, they meet (the counts are all it takes
  // to place the parts), then every thread copies its part
  std::atomic<int> counted{0};
  std::atomic<bool> too_big{false};
  auto both = [&](int t) {
    count(t);
    counted.fetch_add(1, std::memory_order_acq_rel);
    while (counted.load(std::memory_order_acquire) < T) std::this_thread::yield();
    uint64_t before = 0, all = 0;
    for (int u = 0; u < T; ++u) {
      if (u < t) before += raw[u + 1];
      all += raw[u + 1];
    }
    if (all + 1 > cap) { too_big = true; return; }  // (+1: a final record without EOL gains one)
    start[t] = before;
    copy(t);
  };
  if (T == 1) {
    both(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back(both, t);
    both(0);
    for (auto& x : th) x.join();
  }
  if (too_big.load()) return UINT64_MAX;
  // close the parts up (moves nothing when every part is exactly as long as its input records)
  uint64_t pos = 0;
  for (int t = 0; t < T; ++t) {
    const uint64_t from = start[t];
    if (from != pos && wrote[t]) memmove(out + pos, out + from, wrote[t]);
    pos += wrote[t];
  }
  return pos;
}

}  // namespace bk
