// Host-side file plumbing of bk_run (not part of the ABI): a small thread pool, streamed inputs (plain or
// gzip-family, fastq.rs:82-99,114-124) and ordered outputs (plain or gzip / BGZF / mgzip members, fastq.rs:233-249).
// Codecs are zlib on the host's cores; nothing here touches a read's content.
#pragma once
#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "bk_lz4.h"

namespace bkio {

// ---- thread pool: parallel_for with the caller taking part --------------------------------------------------------
class Pool {
 public:
  explicit Pool(unsigned n) {
    for (unsigned i = 0; i < n; ++i) th_.emplace_back([this] { work(); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  unsigned size() const { return (unsigned)th_.size(); }
  // f(i) for i in [0, n); returns when all are done.  Several callers may run loops at the same time.
  void parallel_for(size_t n, const std::function<void(size_t)>& f) {
    if (n == 0) return;
    if (n == 1 || th_.empty()) {
      for (size_t i = 0; i < n; ++i) f(i);
      return;
    }
    auto job = std::make_shared<Job>();
    job->n = n;
    job->f = &f;
    const size_t helpers = std::min<size_t>(n - 1, th_.size());
    {
      std::lock_guard<std::mutex> g(mu_);
      for (size_t h = 0; h < helpers; ++h) q_.push_back(job);
    }
    cv_.notify_all();
    run(*job);
    std::unique_lock<std::mutex> lk(job->mu);
    job->cv.wait(lk, [&] { return job->done == job->n; });
  }

 private:
  struct Job {
    size_t n = 0;
    const std::function<void(size_t)>* f = nullptr;
    std::atomic<size_t> next{0};
    size_t done = 0;
    std::mutex mu;
    std::condition_variable cv;
  };
  static void run(Job& j) {
    size_t mine = 0;
    for (;;) {
      const size_t i = j.next.fetch_add(1);
      if (i >= j.n) break;
      (*j.f)(i);
      ++mine;
    }
    if (mine) {
      std::lock_guard<std::mutex> g(j.mu);
      j.done += mine;
      if (j.done == j.n) j.cv.notify_all();
    }
  }
  void work() {
    for (;;) {
      std::shared_ptr<Job> j;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
        if (q_.empty()) return;
        j = q_.front();
        q_.pop_front();
      }
      run(*j);
    }
  }
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<std::shared_ptr<Job>> q_;
  bool stop_ = false;
};

// ---- inputs ------------------------------------------------------------------------------------------------------

enum class Codec { Plain, Gzip, Lz4 };

// One input file, read front to back.  fill() delivers the next bytes of the (decompressed) text.
class Input {
 public:
  ~Input() {
    if (zinit_) inflateEnd(&zs_);
    if (fd_ >= 0) close(fd_);
  }
  bool open(const char* path, std::string* err) {
    fd_ = ::open(path, O_RDONLY);
    if (fd_ < 0) {  // the reference panics here (fastq.rs:110); we report and exit 1
      *err = std::string("couldn't open file ") + path;
      return false;
    }
    struct stat st;
    if (fstat(fd_, &st) != 0) { *err = std::string("I/O error: ") + strerror(errno); return false; }
    regular_ = S_ISREG(st.st_mode);
    size_ = regular_ ? (uint64_t)st.st_size : 0;
    // CompressionType::detect (fastq.rs:82-99): magic bytes; anything with the gzip magic (gzip, BGZF, mgzip) is
    // read as a multi-member gzip stream, as the reference does (its BGZF branch is unreachable)
    uint8_t head[4];
    size_t got = 0;
    while (got < sizeof head) {
      const ssize_t r = regular_ ? pread(fd_, head + got, sizeof head - got, (off_t)got) : read(fd_, head + got, sizeof head - got);
      if (r < 0 && errno == EINTR) continue;
      if (r <= 0) break;
      got += (size_t)r;
    }
    if (!regular_) pre_.assign(head, head + got);  // a pipe cannot be rewound: those bytes are delivered first
    if (got >= 2 && head[0] == 0x1f && head[1] == 0x8b) codec_ = Codec::Gzip;
    else if (got >= 4 && head[0] == 0x04 && head[1] == 0x22 && head[2] == 0x4d && head[3] == 0x18) codec_ = Codec::Lz4;
    if (codec_ == Codec::Gzip) {
      memset(&zs_, 0, sizeof zs_);
      if (inflateInit2(&zs_, 15 + 16) != Z_OK) { *err = "I/O error: zlib initialisation failed"; return false; }
      zinit_ = true;
      zbuf_.resize(4 << 20);
    }
    return true;
  }
  Codec codec() const { return codec_; }
  bool regular() const { return regular_; }
  // plain regular files may be read by the caller itself, in slices and in parallel: pread at position() + k, then advance()
  bool direct() const { return codec_ == Codec::Plain && regular_; }
  int fd() const { return fd_; }
  uint64_t position() const { return pos_; }
  uint64_t remaining() const { return size_ - pos_; }
  void advance(uint64_t n) {
    pos_ += n;
    if (pos_ >= size_) eof_ = true;
  }
  uint64_t file_size() const { return size_; }
  bool eof() const { return eof_; }

  // Up to `cap` bytes of text into dst; fewer only at the end of the input.  -1 on error.
  ssize_t fill(uint8_t* dst, size_t cap, Pool* pool, std::string* err) {
    if (eof_ || cap == 0) return 0;
    if (codec_ == Codec::Plain) return fill_plain(dst, cap, pool, err);
    if (codec_ == Codec::Lz4) {  // lz4::Decoder (fastq.rs:116)
      const long r = lz4_.read(dst, cap, [&](uint8_t* b, size_t n) { return (long)raw_read(b, n, err); }, err);
      if (r >= 0 && (size_t)r < cap) eof_ = true;
      return (ssize_t)r;
    }
    return fill_gzip(dst, cap, err);
  }

 private:
  ssize_t raw_read(uint8_t* dst, size_t cap, std::string* err) {  // sequential bytes of the file itself
    size_t got = 0;
    if (!pre_.empty()) {
      got = std::min(cap, pre_.size());
      memcpy(dst, pre_.data(), got);
      pre_.erase(pre_.begin(), pre_.begin() + (long)got);
      if (got == cap) return (ssize_t)got;
    }
    while (got < cap) {
      const ssize_t r = regular_ ? pread(fd_, dst + got, cap - got, (off_t)pos_) : read(fd_, dst + got, cap - got);
      if (r < 0) {
        if (errno == EINTR) continue;
        *err = std::string("I/O error: ") + strerror(errno);
        return -1;
      }
      if (r == 0) break;
      got += (size_t)r;
      pos_ += (uint64_t)r;
    }
    return (ssize_t)got;
  }
  ssize_t fill_plain(uint8_t* dst, size_t cap, Pool* pool, std::string* err) {
    if (!regular_) {
      const ssize_t r = raw_read(dst, cap, err);
      if (r >= 0 && (size_t)r < cap) eof_ = true;
      return r;
    }
    const size_t n = (size_t)std::min<uint64_t>(cap, size_ - pos_);
    // the page cache (or the device) serves several readers at once: slices of >= 2 MiB
    const size_t slice = std::max<size_t>(2 << 20, (n + 15) / 16);
    const size_t ns = (n + slice - 1) / slice;
    std::atomic<int> bad{0};
    const uint64_t base = pos_;
    auto part = [&](size_t s) {
      size_t o = s * slice;
      const size_t e = std::min(n, o + slice);
      while (o < e) {
        const ssize_t r = pread(fd_, dst + o, e - o, (off_t)(base + o));
        if (r < 0 && errno == EINTR) continue;
        if (r <= 0) { bad = r < 0 ? errno : EIO; return; }
        o += (size_t)r;
      }
    };
    if (pool) pool->parallel_for(ns, part);
    else for (size_t s = 0; s < ns; ++s) part(s);
    if (bad.load()) { *err = std::string("I/O error: ") + strerror(bad.load()); return -1; }
    pos_ += n;
    if (pos_ >= size_) eof_ = true;
    return (ssize_t)n;
  }
  ssize_t fill_gzip(uint8_t* dst, size_t cap, std::string* err) {  // flate2::read::MultiGzDecoder (fastq.rs:115)
    zs_.next_out = dst;
    zs_.avail_out = (uInt)std::min<size_t>(cap, 1u << 30);
    const size_t want = zs_.avail_out;
    while (zs_.avail_out > 0) {
      if (zs_.avail_in == 0) {
        const ssize_t r = raw_read(zbuf_.data(), zbuf_.size(), err);
        if (r < 0) return -1;
        if (r == 0) {
          if (in_member_) { *err = "I/O error: unexpected end of gzip stream"; return -1; }
          eof_ = true;
          break;
        }
        zs_.next_in = zbuf_.data();
        zs_.avail_in = (uInt)r;
      }
      if (!in_member_) {  // between members: zero padding is tolerated, as gzip(1) does
        while (zs_.avail_in && *zs_.next_in == 0) { ++zs_.next_in; --zs_.avail_in; }
        if (!zs_.avail_in) continue;
        in_member_ = true;
      }
      const int rc = inflate(&zs_, Z_NO_FLUSH);
      if (rc == Z_STREAM_END) {
        in_member_ = false;
        inflateReset(&zs_);
      } else if (rc != Z_OK && rc != Z_BUF_ERROR) {
        *err = std::string("I/O error: corrupt gzip stream") + (zs_.msg ? std::string(" (") + zs_.msg + ")" : std::string());
        return -1;
      }
    }
    return (ssize_t)(want - zs_.avail_out);
  }

  int fd_ = -1;
  bool regular_ = true, eof_ = false;
  uint64_t size_ = 0, pos_ = 0;
  Codec codec_ = Codec::Plain;
  std::vector<uint8_t> pre_;
  z_stream zs_;
  bool zinit_ = false, in_member_ = false;
  std::vector<uint8_t> zbuf_;
  bklz4::Decoder lz4_;
};

// ---- outputs -----------------------------------------------------------------------------------------------------

// What --gz / --bgz / --mgz select (fastq.rs:71-79; --bgz gives Mgzip and --mgz gives Bgzf there, SURVEY F15, and so
// they do here: bk_run maps the flags).
enum class OutFormat { Plain = 0, Gzip = 1, Bgzf = 2, Mgzip = 3, Lz4 = 4 };

inline void put16(std::vector<uint8_t>& v, uint32_t x) { v.push_back((uint8_t)x); v.push_back((uint8_t)(x >> 8)); }
inline void put32(std::vector<uint8_t>& v, uint32_t x) { put16(v, x); put16(v, x >> 16); }

// One gzip member holding src[0, n): plain gzip, a BGZF block (extra subfield 'BC', total size - 1; n <= 0xff00) or
// an mgzip block (extra subfield 'IG', total size).  Level 6 = flate2's Compression::default().
inline bool gzip_member(const uint8_t* src, size_t n, OutFormat fmt, std::vector<uint8_t>* out) {
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (deflateInit2(&zs, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
  const size_t bound = deflateBound(&zs, (uLong)n);
  const size_t h0 = out->size();
  const bool extra = fmt == OutFormat::Bgzf || fmt == OutFormat::Mgzip;
  out->push_back(0x1f); out->push_back(0x8b); out->push_back(8); out->push_back(extra ? 4 : 0);
  put32(*out, 0);                       // mtime
  out->push_back(0); out->push_back(0xff);  // xfl, os = unknown
  size_t size_at = 0;
  if (fmt == OutFormat::Bgzf) {
    put16(*out, 6); out->push_back('B'); out->push_back('C'); put16(*out, 2);
    size_at = out->size(); put16(*out, 0);
  } else if (fmt == OutFormat::Mgzip) {
    put16(*out, 8); out->push_back('I'); out->push_back('G'); put16(*out, 4);
    size_at = out->size(); put32(*out, 0);
  }
  const size_t body = out->size();
  out->resize(body + bound);
  zs.next_in = const_cast<Bytef*>(src);
  zs.avail_in = (uInt)n;
  zs.next_out = out->data() + body;
  zs.avail_out = (uInt)bound;
  const int rc = deflate(&zs, Z_FINISH);
  const size_t clen = bound - zs.avail_out;
  deflateEnd(&zs);
  if (rc != Z_STREAM_END) return false;
  out->resize(body + clen);
  put32(*out, (uint32_t)crc32(crc32(0L, Z_NULL, 0), src, (uInt)n));
  put32(*out, (uint32_t)n);
  const size_t total = out->size() - h0;
  if (fmt == OutFormat::Bgzf) {
    if (total - 1 > 0xffff) return false;
    (*out)[size_at] = (uint8_t)(total - 1); (*out)[size_at + 1] = (uint8_t)((total - 1) >> 8);
  } else if (fmt == OutFormat::Mgzip) {
    for (int i = 0; i < 4; ++i) (*out)[size_at + i] = (uint8_t)(total >> (8 * i));
  }
  return true;
}

// the text of one batch as a run of members, compressed on the pool; `pieces` in order
inline bool compress_batch(const uint8_t* src, size_t n, OutFormat fmt, Pool* pool, std::vector<std::vector<uint8_t>>* pieces) {
  if (fmt == OutFormat::Lz4) {  // blocks of the file's one frame (bk_run writes its header and end mark)
    const size_t nb = (n + bklz4::kBlock - 1) / bklz4::kBlock;
    pieces->assign(nb, std::vector<uint8_t>());
    auto task = [&](size_t t) {
      const size_t b = t * bklz4::kBlock;
      bklz4::frame_block(src + b, std::min(bklz4::kBlock, n - b), &(*pieces)[t]);
    };
    if (pool) pool->parallel_for(nb, task);
    else for (size_t t = 0; t < nb; ++t) task(t);
    return true;
  }
  // BGZF blocks hold at most 64 KiB; gzip / mgzip members are sized for parallelism, not for the format
  const size_t unit = fmt == OutFormat::Bgzf ? 0xff00 : (size_t)1 << 20;
  const size_t group = fmt == OutFormat::Bgzf ? 16 : 1;  // members per task
  const size_t nunits = (n + unit - 1) / unit;
  const size_t ntasks = (nunits + group - 1) / group;
  pieces->assign(ntasks, std::vector<uint8_t>());
  std::atomic<bool> ok{true};
  auto task = [&](size_t t) {
    std::vector<uint8_t>& out = (*pieces)[t];
    for (size_t u = t * group; u < std::min(nunits, (t + 1) * group); ++u) {
      const size_t b = u * unit, e = std::min(n, b + unit);
      if (!gzip_member(src + b, e - b, fmt, &out)) ok = false;
    }
  };
  if (pool) pool->parallel_for(ntasks, task);
  else for (size_t t = 0; t < ntasks; ++t) task(t);
  return ok.load();
}

inline bool pwrite_all(int fd, const uint8_t* p, size_t n, uint64_t off, int* err_no) {
  while (n) {
    const ssize_t w = pwrite(fd, p, n, (off_t)off);
    if (w < 0) {
      if (errno == EINTR) continue;
      *err_no = errno;
      return false;
    }
    p += w; n -= (size_t)w; off += (uint64_t)w;
  }
  return true;
}

}  // namespace bkio
