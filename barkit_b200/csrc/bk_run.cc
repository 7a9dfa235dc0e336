// File -> file runtime: the replacement of barkit-extract/src/run.rs (`run::run`, run.rs:10-60 and the two loops
// process_single_end_fastq :83-146 / process_pair_end_fastq :198-276) on top of the batch ABI.  Host side only, and
// it never looks inside a read:
//
//   reader   one thread streams the inputs (plain: parallel pread; gzip family: zlib, one thread per mate) straight
//            into pinned batch buffers, counts line ends to cut each batch at a record boundary common to both
//            mates, and carries the rest over into the next batch.  Record offsets, the '@' / '+' / length checks
//            (seq_io's reader, fastq.rs:166-172) and everything per read happen on the device (bk_submit_text).
//   workers  one thread per GPU, own bk_ctx, two batches in flight (bk_submit_text / bk_wait).
//   writers  finished batches get their file offsets in batch order (a prefix sum over output lengths, so 1 / 2 / 4 /
//            8-GPU runs write identical files, run.rs:79,194) and are then written concurrently with pwrite;
//            compressed outputs are deflated per batch on the pool first.
//
// No collective and no peer traffic: batches are independent.
#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/barkit_b200.h"
#include "bk_io.h"

namespace {

std::string io_error(const std::string& m) { return "I/O error: " + m; }  // error.rs:15-16

std::string human_duration(double secs) {  // indicatif::HumanDuration (logger.rs:108-110)
  struct U { double s; const char* one; const char* many; };
  static const U units[] = {{365 * 86400.0, "year", "years"}, {7 * 86400.0, "week", "weeks"},
                            {86400.0, "day", "days"},         {3600.0, "hour", "hours"},
                            {60.0, "minute", "minutes"},      {1.0, "second", "seconds"}};
  for (const U& u : units) {
    if (secs >= 1.5 * u.s || u.s == 1.0) {
      long v = (long)(secs / u.s + 0.5);
      char buf[64];
      snprintf(buf, sizeof buf, "%ld %s", v, v == 1 ? u.one : u.many);
      return buf;
    }
  }
  return "0 seconds";
}

struct Logger {  // logger.rs:10-114 (the progress bar itself is not drawn)
  int current = 0, total = 3;
  bool quiet = false;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void message(const char* text) {
    if (current < total) {
      ++current;
      if (!quiet) { printf("[%d/%d] %s\n", current, total, text); fflush(stdout); }
    } else {
      fprintf(stderr, "Warning: Current step exceeds total steps.\n");
    }
  }
  void final_message() {
    if (!quiet) {
      double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      printf("\xE2\x9C\xA8  Done in %s\n", human_duration(s).c_str());
      fflush(stdout);
    }
  }
};

// FastqWriter::new (fastq.rs:210-257): --force logic
int open_output(const char* path, bool force, std::string* err) {
  struct stat st;
  if (stat(path, &st) == 0 && !force) {
    *err = io_error(std::string(path) + " is already existed, use --force to override");  // fastq.rs:219-225
    return -1;
  }
  int fd = force ? open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644) : open(path, O_WRONLY | O_CREAT | O_EXCL, 0644);
  if (fd < 0) *err = io_error(strerror(errno));
  return fd;
}

bool write_all(int fd, const uint8_t* p, size_t n) {
  while (n) {
    ssize_t w = write(fd, p, n);
    if (w < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    p += w;
    n -= (size_t)w;
  }
  return true;
}

void set_err(const std::string& m, char* err, size_t cap) {
  if (!err || !cap) return;
  size_t k = std::min(m.size(), cap - 1);
  memcpy(err, m.data(), k);
  err[k] = 0;
}


inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

size_t count_nl(const uint8_t* p, size_t n) {  // (the plain loop vectorises: compare + horizontal add)
  size_t total = 0;
  while (n) {
    const size_t k = std::min<size_t>(n, 1 << 15);
    unsigned c = 0;
    for (size_t i = 0; i < k; ++i) c += p[i] == '\n';
    total += c;
    p += k;
    n -= k;
  }
  return total;
}

// one batch on its way through the pipeline, with its pinned buffers
struct Set {
  uint8_t* text[2] = {nullptr, nullptr};
  uint8_t* out[2] = {nullptr, nullptr};
  size_t tcap[2] = {0, 0}, ocap[2] = {0, 0};
  size_t len[2] = {0, 0};   // bytes of text of this batch (whole records)
  uint32_t n = 0;           // records (pairs)
  size_t batch = 0;         // index in input order
  bk_result res;
  std::vector<std::vector<uint8_t>> pieces[2];  // compressed outputs
  uint64_t woff[2] = {0, 0};  // file offsets, assigned in batch order
  uint64_t wlen[2] = {0, 0};
  std::atomic<int> writes_left{0};
};

template <class T>
class Queue {  // blocking FIFO; close() lets the consumers drain it and leave, abort() wakes everybody at once
 public:
  void push(T v) {
    {
      std::lock_guard<std::mutex> g(mu_);
      q_.push_back(v);
    }
    cv_.notify_one();
  }
  bool pop(T* v) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return !q_.empty() || closed_ || aborted_; });
    if (aborted_ || q_.empty()) return false;
    *v = q_.front();
    q_.pop_front();
    return true;
  }
  void close() {
    {
      std::lock_guard<std::mutex> g(mu_);
      closed_ = true;
    }
    cv_.notify_all();
  }
  void abort() {
    {
      std::lock_guard<std::mutex> g(mu_);
      aborted_ = true;
    }
    cv_.notify_all();
  }

 private:
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<T> q_;
  bool closed_ = false, aborted_ = false;
};

struct Output {
  int fd = -1;
  bool seekable = true;
  uint64_t pos = 0;    // next free offset (assigned in batch order)
};

}  // namespace

extern "C" int bk_run(const bk_run_args* a, char* err, size_t errcap) {
  if (!a || !a->fq1 || !a->out_fq1) { set_err("missing arguments", err, errcap); return BK_E_ARG; }
  // run.rs:26-59: argument shapes
  const bool pe = a->fq2 && a->out_fq2;
  const bool se = !a->fq2 && !a->out_fq2 && a->pattern1 && !a->pattern2;
  if (!pe && !se) {
    fprintf(stderr,
            "Invalid arguments provided. Please ensure that you have provided the correct combination of input "
            "files and patterns.\n");  // run.rs:56-58: message, normal return
    return BK_OK;
  }
  // a->compression: 0 none, 1 --gz, 2 --bgz, 3 --mgz, 4 --lz4 (src/main.rs:14-19).  CompressionType::select
  // (fastq.rs:71-79) maps --bgz to Mgzip and --mgz to Bgzf; the swap is kept.  --lz4: one LZ4 frame.
  bkio::OutFormat fmt = bkio::OutFormat::Plain;
  switch (a->compression) {
    case 0: break;
    case 1: fmt = bkio::OutFormat::Gzip; break;
    case 2: fmt = bkio::OutFormat::Mgzip; break;
    case 3: fmt = bkio::OutFormat::Bgzf; break;
    case 4: fmt = bkio::OutFormat::Lz4; break;
    default:
      set_err("unknown compression selector", err, errcap);
      return BK_E_ARG;
  }
  std::string e;
  Logger log;
  log.quiet = a->quiet != 0;
  log.message("Estimating reads count...");  // run.rs:97 / :215 (the count only feeds the progress bar, which is not drawn)
  const bool dbg = a->verbose != 0;
  const double t_begin = now_s();
  const int nm = pe ? 2 : 1;

  // the CUDA contexts come up while the host opens its files
  std::thread warm([&] {
    const int nd = bk_device_count();
    const int ng = a->n_gpus > 0 ? std::min(a->n_gpus, nd) : std::min(nd, 1);  // (further devices come up in their workers)
    for (int d = 0; d < ng; ++d) bk_warmup(d);
  });
  struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } warm_join{warm};

  bkio::Input in[2];
  if (!in[0].open(a->fq1, &e)) { set_err(e, err, errcap); return BK_E_IO; }
  if (pe && !in[1].open(a->fq2, &e)) { set_err(e, err, errcap); return BK_E_IO; }
  Output outp[2];
  outp[0].fd = open_output(a->out_fq1, a->force != 0, &e);  // run.rs:106-111 / :225-229
  if (outp[0].fd < 0) { set_err(e, err, errcap); return BK_E_IO; }
  if (pe) {
    outp[1].fd = open_output(a->out_fq2, a->force != 0, &e);
    if (outp[1].fd < 0) { close(outp[0].fd); set_err(e, err, errcap); return BK_E_IO; }
  }
  for (int m = 0; m < nm; ++m) {
    struct stat st;
    outp[m].seekable = fstat(outp[m].fd, &st) == 0 && S_ISREG(st.st_mode);
  }
  auto close_outputs = [&]() { for (int m = 0; m < nm; ++m) close(outp[m].fd); };
  if (fmt == bkio::OutFormat::Lz4) {  // one frame per file: its header now, blocks per batch, the end mark last
    std::vector<uint8_t> head;
    bklz4::frame_header(&head);
    for (int m = 0; m < nm; ++m) {
      if (!write_all(outp[m].fd, head.data(), head.size())) { close_outputs(); set_err(io_error(strerror(errno)), err, errcap); return BK_E_IO; }
      outp[m].pos = head.size();
    }
  }

  log.message("Parsing barcode patterns...");  // run.rs:113 / :231
  bk_program *p1 = nullptr, *p2 = nullptr;
  char msg[1024];
  int st = BK_OK;
  if (a->pattern1 && !(p1 = bk_compile(a->pattern1, a->max_error, &st, msg, sizeof msg))) {
    close_outputs(); set_err(msg, err, errcap); return st;
  }
  if (a->pattern2 && !(p2 = bk_compile(a->pattern2, a->max_error, &st, msg, sizeof msg))) {
    bk_program_destroy(p1); close_outputs(); set_err(msg, err, errcap); return st;
  }
  auto free_programs = [&]() { bk_program_destroy(p1); bk_program_destroy(p2); };
  if (pe && !p1 && !p2) {  // clap: at least one pattern (src/lib.rs:71)
    free_programs(); close_outputs();
    set_err("at least one of --pattern1 / --pattern2 is required", err, errcap);
    return BK_E_ARG;
  }

  log.message("Extracting barcodes from reads...");  // run.rs:120 / :247

  if (warm.joinable()) warm.join();
  const double t_warm = now_s() - t_begin;
  const int ndev = bk_device_count();
  if (ndev <= 0) {
    free_programs(); close_outputs();
    set_err("no CUDA device available; this build has no CPU path", err, errcap);
    return BK_E_CUDA;
  }
  // Batch size per mate: -m MiB when given, else 32 MiB: large enough for the kernels and the bus, small enough to
  // pipeline and to keep the pinned pools (and their allocation time) small.
  size_t batch_bytes = a->max_memory_mb ? a->max_memory_mb << 20 : (size_t)32 << 20;
  batch_bytes = std::min<size_t>(std::max<size_t>(batch_bytes, 1 << 20), (size_t)2 << 30);
  // GPUs: as many as asked for; otherwise one per 2 GiB of plain input per mate (one GPU streams a file faster than
  // most file systems deliver it, and every further context costs its start-up), one for compressed or piped input.
  int ngpu = 1;
  if (a->n_gpus > 0) ngpu = std::min(a->n_gpus, ndev);
  else if (in[0].codec() == bkio::Codec::Plain && in[0].regular())
    ngpu = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)ndev, in[0].file_size() >> 31));

  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const unsigned host_threads = std::min(64u, std::max<unsigned>((unsigned)a->threads, hw));
  bkio::Pool pool(host_threads > 1 ? host_threads - 1 : 0);

  std::atomic<int> status{BK_OK};
  std::string first_error;
  std::mutex emu;
  Queue<Set*> free_sets, ready, tasks;
  Queue<Set*> wq[2];                    // per output file, in batch order
  auto fail = [&](int code, const std::string& m) {
    {
      std::lock_guard<std::mutex> g(emu);
      if (status.load() == BK_OK) { first_error = m; status = code; }
    }
    free_sets.abort(); ready.abort(); tasks.abort(); wq[0].abort(); wq[1].abort();
  };

  const int nsets = 2 * ngpu + 3;  // per GPU two on the device; one being read, two on their way out
  std::vector<Set> sets((size_t)nsets);
  for (auto& s : sets) free_sets.push(&s);
  double t_alloc = 0;
  std::mutex amu;
  auto ensure_host = [&](uint8_t** p, size_t* cap, size_t need) {
    if (*cap >= need) return true;
    const double t0 = now_s();
    if (*p) bk_host_free(*p);
    const size_t want = need + need / 16 + 4096;
    *p = (uint8_t*)bk_host_alloc(want);
    *cap = *p ? want : 0;
    {
      std::lock_guard<std::mutex> g(amu);
      t_alloc += now_s() - t0;
    }
    return *p != nullptr;
  };

  // ---- reader ------------------------------------------------------------------------------------------------------
  double t_read = 0, t_count = 0, t_free_wait = 0, t_first_ready = 0, t_reader_done = 0, t_workers_done = 0;
  size_t nbatches = 0;
  uint64_t nrec_total = 0;
  std::thread reader([&] {
    std::vector<uint8_t> carry[2];
    size_t target[2] = {batch_bytes, batch_bytes};
    for (size_t bi = 0;; ++bi) {
      Set* s = nullptr;
      const double tf = now_s();
      if (!free_sets.pop(&s)) return;
      t_free_wait += now_s() - tf;
      size_t len[2] = {0, 0};
      for (int m = 0; m < nm; ++m) {
        const size_t need = std::max(target[m], carry[m].size()) + 64;
        if (!ensure_host(&s->text[m], &s->tcap[m], std::max(need, batch_bytes + 64))) { fail(BK_E_CAPACITY, "pinned host allocation failed"); return; }
        if (!carry[m].empty()) memcpy(s->text[m], carry[m].data(), carry[m].size());
        len[m] = carry[m].size();
        carry[m].clear();
      }
      const double tr = now_s();
      // Segments of the batch text: the carried-over front, then the new bytes in 1 MiB pieces.  Plain files are read
      // piece by piece on the pool, each piece's line ends counted while it is still in the core's cache; other
      // inputs (gzip, pipes) are filled first, one thread per mate, and counted afterwards.
      struct Seg { int m; size_t b, e; uint64_t fpos; bool read; uint32_t nl; };
      std::vector<Seg> segs;
      constexpr size_t kSeg = 1 << 20;
      std::string ferr[2];
      bool fok[2] = {true, true};
      auto goal_of = [&](int m) { return std::min(std::max(target[m], len[m]), s->tcap[m] - 64); };
      auto fill = [&](int m) {
        const size_t goal = goal_of(m);
        while (len[m] < goal && !in[m].eof()) {
          const ssize_t r = in[m].fill(s->text[m] + len[m], goal - len[m], nullptr, &ferr[m]);
          if (r < 0) { fok[m] = false; return; }
          len[m] += (size_t)r;
        }
      };
      size_t old_len[2] = {len[0], len[1]};
      if (pe && !in[0].direct() && !in[1].direct()) {
        std::thread t2(fill, 1);
        fill(0);
        t2.join();
      } else {
        for (int m = 0; m < nm; ++m)
          if (!in[m].direct()) fill(m);
      }
      for (int m = 0; m < nm; ++m) {
        if (!fok[m]) { fail(BK_E_IO, ferr[m]); return; }
        if (old_len[m]) segs.push_back({m, 0, old_len[m], 0, false, 0});
        if (in[m].direct()) {
          const size_t goal = goal_of(m);
          const size_t n = goal > len[m] ? (size_t)std::min<uint64_t>(goal - len[m], in[m].remaining()) : 0;
          for (size_t o = 0; o < n; o += kSeg)
            segs.push_back({m, len[m] + o, len[m] + std::min(n, o + kSeg), in[m].position() + o, true, 0});
          len[m] += n;
          in[m].advance(n);
        } else {
          for (size_t o = old_len[m]; o < len[m]; o += kSeg) segs.push_back({m, o, std::min(len[m], o + kSeg), 0, false, 0});
        }
      }
      std::atomic<int> rbad{0};
      pool.parallel_for(segs.size(), [&](size_t i) {
        Seg& g = segs[i];
        uint8_t* t = s->text[g.m];
        if (g.read) {
          size_t o = g.b;
          while (o < g.e) {
            const ssize_t r = pread(in[g.m].fd(), t + o, g.e - o, (off_t)(g.fpos + (o - g.b)));
            if (r < 0 && errno == EINTR) continue;
            if (r <= 0) { rbad = r < 0 ? errno : EIO; return; }
            o += (size_t)r;
          }
        }
        g.nl = (uint32_t)count_nl(t + g.b, g.e - g.b);
      });
      if (rbad.load()) { fail(BK_E_IO, io_error(strerror(rbad.load()))); return; }
      t_read += now_s() - tr;
      const double tc = now_s();
      for (int m = 0; m < nm; ++m) {
        if (!in[m].eof()) continue;
        // blank lines after the last record are not records (seq_io accepts them): cut them off, recount the tail
        uint8_t* t = s->text[m];
        size_t& L = len[m];
        const size_t L0 = L;
        while (L && t[L - 1] == '\n') {
          if (L == 1 || t[L - 2] == '\n') L -= 1;
          else if (t[L - 2] == '\r' && (L == 2 || t[L - 3] == '\n')) L -= 2;
          else break;
        }
        if (L != L0)
          for (Seg& g : segs)
            if (g.m == m && g.e > L) {
              g.e = L;
              g.b = std::min(g.b, L);
              g.nl = (uint32_t)count_nl(t + g.b, g.e - g.b);
            }
      }
      // per mate: segment bounds and running line-end counts
      std::vector<size_t> sb[2];
      std::vector<uint64_t> cnt[2];
      for (int m = 0; m < nm; ++m) { sb[m].push_back(0); cnt[m].push_back(0); }
      for (const Seg& g : segs) {
        sb[g.m].push_back(g.e);
        cnt[g.m].push_back(cnt[g.m].back() + g.nl);
      }
      uint64_t avail[2] = {UINT64_MAX, UINT64_MAX}, lines[2] = {0, 0}, nl[2] = {0, 0};
      bool ended = false;
      for (int m = 0; m < nm; ++m) {
        nl[m] = cnt[m].back();
        const bool open_line = len[m] && s->text[m][len[m] - 1] != '\n';
        lines[m] = nl[m] + (in[m].eof() && open_line ? 1 : 0);  // a last record without EOL is a record (seq_io)
        avail[m] = lines[m] / 4;
        if (in[m].eof() && avail[m] == 0) {
          if (lines[m] != 0) { fail(BK_E_FORMAT, "malformed FASTQ: the input ends inside a record"); return; }
          ended = true;  // (in paired mode the shorter input ends the run: zip, run.rs:172-173)
        }
      }
      uint64_t n = std::min(avail[0], pe ? avail[1] : avail[0]);
      n = std::min<uint64_t>(n, 1u << 30);
      if (ended || n == 0) {
        if (!ended) { fail(BK_E_CAPACITY, "a record is longer than the batch buffer (raise -m)"); return; }
        free_sets.push(s);
        break;
      }
      for (int m = 0; m < nm; ++m) {
        // end of record n = one past line end number 4n (or the end of the text for an unterminated last record)
        size_t end = len[m];
        const uint64_t want = 4 * n;
        if (want <= nl[m]) {
          const size_t sl = (size_t)(std::lower_bound(cnt[m].begin(), cnt[m].end(), want) - cnt[m].begin());  // first running count >= want
          // the line end lies in segment sl - 1
          size_t pos = sb[m][sl - 1];
          uint64_t seen = cnt[m][sl - 1];
          const uint8_t* t = s->text[m];
          while (seen < want) {
            const uint8_t* q = (const uint8_t*)memchr(t + pos, '\n', len[m] - pos);
            pos = (size_t)(q - t) + 1;
            ++seen;
          }
          end = pos;
        }
        if (in[m].eof() && avail[m] == n && lines[m] % 4 != 0) {  // what follows the last whole record is a fragment
          fail(BK_E_FORMAT, "malformed FASTQ: the input ends inside a record");
          return;
        }
        carry[m].assign(s->text[m] + end, s->text[m] + len[m]);
        s->len[m] = end;
      }
      // next batch: read so that both mates bring about the same number of records
      if (pe) {
        const double a0 = (double)s->len[0] / (double)n, a1 = (double)s->len[1] / (double)n;
        const double nt = std::min((double)batch_bytes / a0, (double)batch_bytes / a1);
        target[0] = std::min(batch_bytes, (size_t)(nt * a0 * 1.02) + 65536);
        target[1] = std::min(batch_bytes, (size_t)(nt * a1 * 1.02) + 65536);
      }
      t_count += now_s() - tc;
      s->n = (uint32_t)n;
      s->batch = bi;
      if (!nbatches) t_first_ready = now_s() - t_begin;
      ++nbatches;
      nrec_total += n;
      ready.push(s);
    }
    t_reader_done = now_s() - t_begin;
    ready.close();
  });

  // ---- sequencer + writers -------------------------------------------------------------------------------------------
  // Finished batches get their file offsets in batch order; each output file then has ONE writer thread that takes
  // the batches in that order.  (write(2) holds the file's lock for the whole call, so several threads writing one
  // file get one thread's rate and pay for the hand-overs: tools/fs_peak.py, 4 GB/s per tmpfs file from 1 or 16
  // threads, 9 GB/s into two files.  A shared mapping filled by all threads reaches 7.5 / 14.6 GB/s in isolation,
  // but inside this process its page faults and unmaps fight the pinned allocations and the DMA set-up for the
  // address-space lock: measured slower end to end, and "disk full" becomes SIGBUS.)
  std::mutex qmu;                       // guards pending / turn / positions / outstanding
  std::map<size_t, Set*> pending;       // finished (and compressed) batches that wait for their turn
  size_t turn = 0;
  size_t outstanding = 0;               // batches between bk_wait and the end of their last write
  bool workers_done = false;
  double t_write[2] = {0, 0}, t_comp = 0;
  std::mutex tmu;
  auto close_if_drained = [&] {         // (qmu held)
    if (workers_done && outstanding == 0) { tasks.close(); wq[0].close(); wq[1].close(); }
  };
  auto sequence = [&](Set* s) {
    std::lock_guard<std::mutex> g(qmu);
    pending[s->batch] = s;
    for (auto it = pending.find(turn); it != pending.end(); it = pending.find(turn)) {
      Set* x = it->second;
      pending.erase(it);
      for (int m = 0; m < nm; ++m) {
        x->woff[m] = outp[m].pos;
        outp[m].pos += x->wlen[m];
      }
      ++turn;
      x->writes_left = nm;
      for (int m = 0; m < nm; ++m) wq[m].push(x);
    }
  };
  auto file_writer = [&](int m) {
    Set* s = nullptr;
    while (wq[m].pop(&s)) {
      const double t0 = now_s();
      int en = 0;
      bool ok = true;
      auto put = [&](const uint8_t* p, size_t n, uint64_t off) {
        if (outp[m].seekable) return bkio::pwrite_all(outp[m].fd, p, n, off, &en);
        if (!write_all(outp[m].fd, p, n)) { en = errno ? errno : EIO; return false; }
        return true;
      };
      if (fmt == bkio::OutFormat::Plain) {
        ok = put(s->out[m], (size_t)s->wlen[m], s->woff[m]);
      } else {
        uint64_t off = s->woff[m];
        for (auto& pc : s->pieces[m]) {
          if (!(ok = put(pc.data(), pc.size(), off))) break;
          off += pc.size();
        }
        s->pieces[m].clear();
      }
      t_write[m] += now_s() - t0;
      if (!ok) { fail(BK_E_IO, io_error(strerror(en))); return; }
      if (s->writes_left.fetch_sub(1) == 1) {
        {
          std::lock_guard<std::mutex> g(qmu);
          --outstanding;
          close_if_drained();
        }
        free_sets.push(s);
      }
    }
  };
  auto compressor = [&] {  // compressed outputs: deflate a batch (all cores), then it joins the line
    Set* s = nullptr;
    while (tasks.pop(&s)) {
      const double t0 = now_s();
      bool ok = true;
      for (int m = 0; m < nm; ++m) {
        const uint64_t n = m ? s->res.out2_len : s->res.out1_len;
        ok = ok && bkio::compress_batch(s->out[m], (size_t)n, fmt, &pool, &s->pieces[m]);
        uint64_t tot = 0;
        for (auto& pc : s->pieces[m]) tot += pc.size();
        s->wlen[m] = tot;
      }
      {
        std::lock_guard<std::mutex> g(tmu);
        t_comp += now_s() - t0;
      }
      if (!ok) { fail(BK_E_IO, "I/O error: deflate failed"); return; }
      sequence(s);
    }
  };
  auto complete = [&](Set* s) {  // called by a worker after bk_wait
    {
      std::lock_guard<std::mutex> g(qmu);
      ++outstanding;
    }
    if (fmt == bkio::OutFormat::Plain) {
      s->wlen[0] = s->res.out1_len;
      s->wlen[1] = pe ? s->res.out2_len : 0;
      sequence(s);
    } else {
      tasks.push(s);
    }
  };
  std::vector<std::thread> wth;
  for (int m = 0; m < nm; ++m) wth.emplace_back(file_writer, m);
  if (fmt != bkio::OutFormat::Plain)
    for (int i = 0; i < std::max(2, ngpu); ++i) wth.emplace_back(compressor);

  // ---- workers: one per GPU ----------------------------------------------------------------------------------------
  const uint32_t flags = (pe ? BK_FLAG_PAIRED : 0) | (a->skip_trimming ? BK_FLAG_SKIP_TRIM : 0) |
                         (a->rc_barcodes ? BK_FLAG_RC : 0) | (dbg ? BK_FLAG_VERBOSE : 0) | BK_FLAG_DEVICE_BOTH_MATES;
  auto worker = [&](int dev) {
    char emsg[512];
    int cst = BK_OK;
    double t_ctx = now_s(), t_submit = 0, t_wait = 0, t_pop = 0;
    bk_ctx* ctx = bk_ctx_create(dev, p1, p2, flags, &cst, emsg, sizeof emsg);
    if (!ctx) { fail(cst, emsg); return; }
    t_ctx = now_s() - t_ctx;
    auto finish = [&](int dslot, Set* s) -> bool {
      const double t0 = now_s();
      const int rc = bk_wait(ctx, dslot, s->out[0], s->ocap[0], s->out[1], s->ocap[1], &s->res);
      t_wait += now_s() - t0;
      if (rc != BK_OK) { fail(rc, bk_last_error(ctx)); return false; }
      complete(s);
      return true;
    };
    Set* prev = nullptr;
    long k = 0;
    bool ok = true;
    for (;; ++k) {
      Set* s = nullptr;
      const double t0 = now_s();
      if (!ready.pop(&s)) break;
      t_pop += now_s() - t0;
      for (int m = 0; m < nm && ok; ++m)
        ok = ensure_host(&s->out[m], &s->ocap[m], std::max<size_t>(bk_out_bound(ctx, m + 1, s->n, s->len[m]), batch_bytes + batch_bytes / 4));
      if (!ok) { fail(BK_E_CAPACITY, "pinned host allocation failed"); break; }
      const double t1 = now_s();
      const int rc = bk_submit_text(ctx, (int)(k & 1), s->n, s->text[0], s->len[0], pe ? s->text[1] : nullptr, pe ? s->len[1] : 0, 0);
      t_submit += now_s() - t1;
      if (rc != BK_OK) { fail(rc, bk_last_error(ctx)); ok = false; break; }
      if (prev && !(ok = finish((int)((k - 1) & 1), prev))) { prev = nullptr; break; }
      prev = s;
    }
    if (ok && prev && status.load() == BK_OK) finish((int)((k - 1) & 1), prev);
    const double td = now_s();
    bk_ctx_destroy(ctx);
    if (dbg) fprintf(stderr, "[bk_run] gpu %d: context %.3f s, waiting for a batch %.3f s, submit %.3f s, wait %.3f s, context released in %.3f s\n", dev,
                     t_ctx, t_pop, t_submit, t_wait, now_s() - td);
  };
  {
    std::vector<std::thread> th;
    for (int d = 0; d < ngpu; ++d) th.emplace_back(worker, d);
    for (auto& t : th) t.join();
  }
  reader.join();
  t_workers_done = now_s() - t_begin;
  {
    std::lock_guard<std::mutex> g(qmu);
    workers_done = true;
    close_if_drained();
  }
  for (auto& t : wth) t.join();
  const double t_writers_done = now_s() - t_begin;

  // the formats' closing pieces: BGZF ends with an empty block, and a gzip file holds at least one member
  if (status.load() == BK_OK && fmt != bkio::OutFormat::Plain) {
    for (int m = 0; m < nm; ++m) {
      if (fmt == bkio::OutFormat::Lz4 || fmt == bkio::OutFormat::Bgzf || outp[m].pos == 0) {
        std::vector<uint8_t> tail;
        if (fmt == bkio::OutFormat::Lz4) tail.assign(4, 0);
        else bkio::gzip_member(nullptr, 0, fmt, &tail);
        int en = 0;
        const bool ok = outp[m].seekable ? bkio::pwrite_all(outp[m].fd, tail.data(), tail.size(), outp[m].pos, &en)
                                         : write_all(outp[m].fd, tail.data(), tail.size());
        if (!ok) fail(BK_E_IO, io_error(strerror(en ? en : errno)));
        outp[m].pos += tail.size();
      }
    }
  }
  const double t_free0 = now_s();
  for (auto& s : sets)
    for (int m = 0; m < 2; ++m) {
      if (s.text[m]) bk_host_free(s.text[m]);
      if (s.out[m]) bk_host_free(s.out[m]);
    }
  const double t_free = now_s() - t_free0;
  free_programs();
  close_outputs();
  if (dbg)
    fprintf(stderr, "[bk_run] timeline: CUDA up %.3f s, first batch read %.3f s, input read %.3f s, device done %.3f s, "
            "outputs written %.3f s, pinned buffers released in %.3f s\n", t_warm, t_first_ready, t_reader_done, t_workers_done, t_writers_done, t_free);
  if (dbg)
    fprintf(stderr, "[bk_run] total %.3f s, %zu batches, %llu records, %d GPU(s), %u host threads; reader: waiting for a free buffer set "
            "%.3f s, read %.3f s, line ends %.3f s; pinned allocation %.3f s; deflate %.3f s; write %.3f + %.3f s\n", now_s() - t_begin,
            nbatches, (unsigned long long)nrec_total, ngpu, host_threads, t_free_wait, t_read, t_count, t_alloc, t_comp, t_write[0], t_write[1]);
  if (status.load() != BK_OK) { set_err(first_error, err, errcap); return status.load(); }
  log.final_message();
  return BK_OK;
}
