// LZ4 frame format, written here from its specification (lz4_Frame_format.md / lz4_Block_format.md): the image
// has liblz4's runtime but not its headers.  What the reference uses it for: `--lz4` outputs (lz4::EncoderBuilder,
// fastq.rs:244) and inputs whose first bytes are 04 22 4D 18 (lz4::Decoder, fastq.rs:116).  Host-side codec only;
// nothing here touches a read's content.
//
//   frame   = magic 0x184D2204 | FLG | BD | [content size 8] | [dict id 4] | HC | block* | 0x00000000 | [content xxh32]
//   block   = u32 size (bit 31: stored uncompressed) | data | [block xxh32]
//   data    = sequences: token (literal length << 4 | match length - 4), [length bytes], literals, u16 offset, [length bytes]
//
// The encoder writes independent 1 MiB blocks (one pool task each), no checksums; the decoder takes anything the
// format allows (linked or independent blocks, checksums present or not, several frames, skippable frames).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

namespace bklz4 {

inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

inline uint32_t xxh32(const uint8_t* p, size_t n, uint32_t seed) {
  const uint32_t P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
  const uint8_t* const e = p + n;
  uint32_t h;
  if (n >= 16) {
    uint32_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
    do {
      v1 = rotl(v1 + rd32(p) * P2, 13) * P1;
      v2 = rotl(v2 + rd32(p + 4) * P2, 13) * P1;
      v3 = rotl(v3 + rd32(p + 8) * P2, 13) * P1;
      v4 = rotl(v4 + rd32(p + 12) * P2, 13) * P1;
      p += 16;
    } while (p + 16 <= e);
    h = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
  } else {
    h = seed + P5;
  }
  h += (uint32_t)n;
  while (p + 4 <= e) { h = rotl(h + rd32(p) * P3, 17) * P4; p += 4; }
  while (p < e) { h = rotl(h + *p * P5, 11) * P1; ++p; }
  h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16;
  return h;
}

// ---- block codec ---------------------------------------------------------------------------------------------------

// Decodes one block into [op, oend); matches may reach back to `obase` (earlier output of a linked frame).
// -> decoded bytes, or -1 for a malformed block.
inline long decode_block(const uint8_t* ip, size_t n, uint8_t* obase, uint8_t* op, uint8_t* oend) {
  const uint8_t* const iend = ip + n;
  uint8_t* const o0 = op;
  while (ip < iend) {
    const unsigned token = *ip++;
    size_t lit = token >> 4;
    if (lit == 15) {
      unsigned s;
      do {
        if (ip >= iend) return -1;
        s = *ip++;
        lit += s;
      } while (s == 255);
    }
    if ((size_t)(iend - ip) < lit || (size_t)(oend - op) < lit) return -1;
    memcpy(op, ip, lit);
    op += lit;
    ip += lit;
    if (ip >= iend) break;  // the last sequence is literals only
    if (iend - ip < 2) return -1;
    const size_t off = (size_t)ip[0] | ((size_t)ip[1] << 8);
    ip += 2;
    size_t ml = token & 15u;
    if (ml == 15) {
      unsigned s;
      do {
        if (ip >= iend) return -1;
        s = *ip++;
        ml += s;
      } while (s == 255);
    }
    ml += 4;
    if (off == 0 || (size_t)(op - obase) < off || (size_t)(oend - op) < ml) return -1;
    const uint8_t* m = op - off;
    if (off >= ml) memcpy(op, m, ml);
    else for (size_t i = 0; i < ml; ++i) op[i] = m[i];  // overlapping: the match repeats itself
    op += ml;
  }
  return (long)(op - o0);
}

inline size_t block_bound(size_t n) { return n + n / 255 + 16; }

// Greedy single-pass compressor (hash of 4 bytes -> last position), independent block.  -> compressed bytes
// (always fits block_bound(n)).
inline size_t encode_block(const uint8_t* src, size_t n, uint8_t* dst) {
  uint8_t* op = dst;
  size_t anchor = 0;
  auto emit = [&](size_t lit_end, size_t off, size_t ml) {  // literals [anchor, lit_end), then a match (ml >= 4) unless off == 0
    const size_t lit = lit_end - anchor;
    uint8_t* const tok = op++;
    unsigned t = 0;
    if (lit >= 15) {
      t = 15u << 4;
      size_t r = lit - 15;
      for (; r >= 255; r -= 255) *op++ = 255;
      *op++ = (uint8_t)r;
    } else {
      t = (unsigned)lit << 4;
    }
    memcpy(op, src + anchor, lit);
    op += lit;
    if (off) {
      *op++ = (uint8_t)off;
      *op++ = (uint8_t)(off >> 8);
      size_t m = ml - 4;
      if (m >= 15) {
        t |= 15u;
        m -= 15;
        for (; m >= 255; m -= 255) *op++ = 255;
        *op++ = (uint8_t)m;
      } else {
        t |= (unsigned)m;
      }
    }
    *tok = (uint8_t)t;
  };
  if (n >= 13) {
    constexpr int kHashBits = 14;
    std::vector<uint32_t> table((size_t)1 << kHashBits, 0xFFFFFFFFu);
    const size_t mflimit = n - 12, matchlimit = n - 5;  // the format's end-of-block rules
    size_t i = 0, miss = 0;
    while (i < mflimit) {
      const uint32_t v = rd32(src + i);
      const uint32_t h = (v * 2654435761u) >> (32 - kHashBits);
      const uint32_t ref = table[h];
      table[h] = (uint32_t)i;
      if (ref != 0xFFFFFFFFu && i - ref <= 65535 && rd32(src + ref) == v) {
        size_t ml = 4;
        while (i + ml < matchlimit && src[ref + ml] == src[i + ml]) ++ml;
        emit(i, i - ref, ml);
        i += ml;
        anchor = i;
        miss = 0;
      } else {
        i += 1 + (miss++ >> 6);  // incompressible stretches are skipped faster and faster
      }
    }
  }
  emit(n, 0, 0);
  return (size_t)(op - dst);
}

// ---- frame -----------------------------------------------------------------------------------------------------------

constexpr uint32_t kMagic = 0x184D2204u;
constexpr size_t kBlock = 1 << 20;  // BD code 6

inline void frame_header(std::vector<uint8_t>* out) {
  const uint8_t desc[2] = {0x60, 0x60};  // FLG: version 01, independent blocks, no checksums, no sizes; BD: 1 MiB blocks
  out->push_back(0x04); out->push_back(0x22); out->push_back(0x4D); out->push_back(0x18);
  out->push_back(desc[0]);
  out->push_back(desc[1]);
  out->push_back((uint8_t)(xxh32(desc, 2, 0) >> 8));
}

// one block of the frame: size word + data (stored when it does not shrink)
inline void frame_block(const uint8_t* src, size_t n, std::vector<uint8_t>* out) {
  const size_t h0 = out->size();
  out->resize(h0 + 4 + block_bound(n));
  size_t c = encode_block(src, n, out->data() + h0 + 4);
  uint32_t word = (uint32_t)c;
  if (c >= n) {
    memcpy(out->data() + h0 + 4, src, n);
    c = n;
    word = (uint32_t)n | 0x80000000u;
  }
  for (int i = 0; i < 4; ++i) (*out)[h0 + i] = (uint8_t)(word >> (8 * i));
  out->resize(h0 + 4 + c);
}

// Streaming frame decoder: feed() it the file's bytes through `more`, take text out with read().
class Decoder {
 public:
  // more(buf, cap) -> bytes read (0 at the end of the file, < 0 on error)
  template <class More>
  long read(uint8_t* dst, size_t cap, More&& more, std::string* err) {
    size_t got = 0;
    while (got < cap) {
      if (pend_ < pend_end_) {
        const size_t k = std::min(cap - got, pend_end_ - pend_);
        memcpy(dst + got, win_.data() + pend_, k);
        pend_ += k;
        got += k;
        continue;
      }
      if (done_) break;
      const int rc = next_block(more, err);
      if (rc < 0) return -1;
      if (rc == 0) { done_ = true; break; }
    }
    return (long)got;
  }

 private:
  template <class More>
  bool need(size_t n, More& more, bool* eof, std::string* err) {  // make n bytes available at in_[ipos_..]
    *eof = false;
    while (in_.size() - ipos_ < n) {
      if (ipos_ > (1u << 20)) { in_.erase(in_.begin(), in_.begin() + (long)ipos_); ipos_ = 0; }
      const size_t old = in_.size();
      in_.resize(old + (1u << 20));
      const long r = more(in_.data() + old, (size_t)1 << 20);
      if (r < 0) { in_.resize(old); return false; }
      in_.resize(old + (size_t)r);
      if (r == 0) { *eof = true; return false; }
    }
    (void)err;
    return true;
  }
  // -> 1 a block was decoded into the window, 0 end of input, -1 error
  template <class More>
  int next_block(More& more, std::string* err) {
    bool eof = false;
    for (;;) {
      if (!in_frame_) {
        if (!need(4, more, &eof, err)) {
          if (eof && in_.size() == ipos_) return 0;  // clean end between frames
          *err = eof ? "I/O error: truncated lz4 stream" : *err;
          return -1;
        }
        const uint32_t magic = rd32(&in_[ipos_]);
        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {  // skippable frame
          if (!need(8, more, &eof, err)) { *err = "I/O error: truncated lz4 stream"; return -1; }
          const size_t len = rd32(&in_[ipos_ + 4]);
          if (!need(8 + len, more, &eof, err)) { *err = "I/O error: truncated lz4 stream"; return -1; }
          ipos_ += 8 + len;
          continue;
        }
        if (magic != kMagic) { *err = "I/O error: not an lz4 frame"; return -1; }
        if (!need(7, more, &eof, err)) { *err = "I/O error: truncated lz4 stream"; return -1; }
        const unsigned flg = in_[ipos_ + 4], bd = in_[ipos_ + 5];
        if ((flg >> 6) != 1) { *err = "I/O error: unsupported lz4 frame version"; return -1; }
        blk_sum_ = (flg >> 4) & 1;
        content_sum_ = (flg >> 2) & 1;
        const size_t extra = ((flg >> 3) & 1 ? 8 : 0) + ((flg & 1) ? 4 : 0);
        const unsigned code = (bd >> 4) & 7;
        if (code < 4) { *err = "I/O error: bad lz4 block size"; return -1; }
        bmax_ = (size_t)1 << (8 + 2 * code);  // 4: 64 KiB ... 7: 4 MiB
        if (!need(7 + extra, more, &eof, err)) { *err = "I/O error: truncated lz4 stream"; return -1; }
        if ((uint8_t)(xxh32(&in_[ipos_ + 4], 2 + extra, 0) >> 8) != in_[ipos_ + 6 + extra]) {
          *err = "I/O error: lz4 frame header checksum mismatch";
          return -1;
        }
        ipos_ += 7 + extra;
        in_frame_ = true;
        hist_ = pend_ = pend_end_ = 0;  // a new frame starts without history (everything decoded so far has been handed out)
      }
      if (!need(4, more, &eof, err)) { *err = "I/O error: truncated lz4 stream"; return -1; }
      const uint32_t word = rd32(&in_[ipos_]);
      ipos_ += 4;
      if (word == 0) {  // end mark
        if (content_sum_) {
          if (!need(4, more, &eof, err)) { *err = "I/O error: truncated lz4 stream"; return -1; }
          ipos_ += 4;
        }
        in_frame_ = false;
        continue;
      }
      const bool stored = (word >> 31) != 0;
      const size_t n = word & 0x7FFFFFFFu;
      if (n > bmax_) { *err = "I/O error: corrupt lz4 stream (block larger than the frame allows)"; return -1; }
      if (!need(n + (blk_sum_ ? 4 : 0), more, &eof, err)) { *err = "I/O error: truncated lz4 stream"; return -1; }
      // window = [history (<= 64 KiB) | this block's output]
      if (win_.size() < kHist + bmax_) win_.resize(kHist + bmax_);
      if (pend_end_ > kHist) {  // keep the last 64 KiB of what was produced so far as the next block's history
        const size_t have = pend_end_ - (kHist - hist_);
        const size_t keep = std::min(have, kHist);
        memmove(win_.data() + kHist - keep, win_.data() + pend_end_ - keep, keep);
        hist_ = keep;
      }
      uint8_t* const op = win_.data() + kHist;
      long d;
      if (stored) {
        memcpy(op, &in_[ipos_], n);
        d = (long)n;
      } else {
        d = decode_block(&in_[ipos_], n, win_.data() + kHist - hist_, op, op + bmax_);
      }
      if (d < 0) { *err = "I/O error: corrupt lz4 stream"; return -1; }
      ipos_ += n + (blk_sum_ ? 4 : 0);
      pend_ = kHist;
      pend_end_ = kHist + (size_t)d;
      return 1;
    }
  }

  static constexpr size_t kHist = 64 << 10;
  std::vector<uint8_t> in_, win_;
  size_t ipos_ = 0, pend_ = 0, pend_end_ = 0, hist_ = 0, bmax_ = 0;
  bool in_frame_ = false, done_ = false, blk_sum_ = false, content_sum_ = false;
};

}  // namespace bklz4
