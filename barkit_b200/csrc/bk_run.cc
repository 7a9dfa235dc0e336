// File -> file runtime: the replacement of barkit-extract/src/run.rs (`run::run`, run.rs:10-60 and
// the two loops process_single_end_fastq :83-146 / process_pair_end_fastq :198-276) on top of the
// batch ABI.  Host side only: memory-mapped plain FASTQ in, parallel record-boundary scan, batches
// dealt to one worker thread per GPU (each with its own bk_ctx, two slots in flight, pinned staging),
// outputs written strictly in batch order so that 1/2/4/8-GPU runs produce identical bytes.
// No collective and no peer traffic: batches are independent.
#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/barkit_b200.h"

namespace {

struct Mapped {
  const uint8_t* p = nullptr;
  size_t len = 0;
  int fd = -1;
  ~Mapped() {
    if (p && len) munmap((void*)p, len);
    if (fd >= 0) close(fd);
  }
};

std::string io_error(const std::string& m) { return "I/O error: " + m; }  // error.rs:15-16

bool map_file(const char* path, Mapped* m, std::string* err) {
  m->fd = open(path, O_RDONLY);
  if (m->fd < 0) {  // the reference panics here (fastq.rs:110); we report and exit 1
    *err = std::string("couldn't open file ") + path;
    return false;
  }
  struct stat st;
  if (fstat(m->fd, &st) != 0) { *err = io_error(strerror(errno)); return false; }
  m->len = (size_t)st.st_size;
  if (m->len == 0) return true;
  void* p = mmap(nullptr, m->len, PROT_READ, MAP_PRIVATE, m->fd, 0);
  if (p == MAP_FAILED) { *err = io_error(strerror(errno)); return false; }
  madvise(p, m->len, MADV_SEQUENTIAL);
  m->p = (const uint8_t*)p;
  return true;
}

// CompressionType::detect (fastq.rs:82-99): compressed inputs are recognised and refused
const char* compressed_kind(const Mapped& m) {
  if (m.len >= 2 && m.p[0] == 0x1f && m.p[1] == 0x8b) return "gzip";
  if (m.len >= 4 && m.p[0] == 0x04 && m.p[1] == 0x22 && m.p[2] == 0x4d && m.p[3] == 0x18) return "lz4";
  return nullptr;
}

// Record starts of a whole file.  FASTQ as seq_io reads it is four lines per record, so record k
// starts at line 4k: count newlines per chunk in parallel, prefix-sum, then emit the starts.
bool record_starts(const Mapped& m, unsigned threads, std::vector<uint64_t>* starts, std::string* err) {
  starts->clear();
  if (m.len == 0) { starts->push_back(0); return true; }
  threads = std::max(1u, threads);
  const size_t chunk = std::max<size_t>((m.len + threads - 1) / threads, 1 << 20);
  const size_t nchunks = (m.len + chunk - 1) / chunk;
  std::vector<uint64_t> nl(nchunks + 1, 0);
  auto count = [&](size_t c) {
    const uint8_t* p = m.p + c * chunk;
    const uint8_t* e = m.p + std::min(m.len, (c + 1) * chunk);
    uint64_t n = 0;
    while (p < e) {
      const void* q = memchr(p, '\n', e - p);
      if (!q) break;
      ++n;
      p = (const uint8_t*)q + 1;
    }
    nl[c + 1] = n;
  };
  {
    std::vector<std::thread> th;
    for (size_t c = 0; c < nchunks; ++c) th.emplace_back(count, c);
    for (auto& t : th) t.join();
  }
  for (size_t c = 0; c < nchunks; ++c) nl[c + 1] += nl[c];
  // a final line without EOL still counts (seq_io accepts a last record without trailing newline)
  const uint64_t data_lines = nl[nchunks] + (m.p[m.len - 1] != '\n' ? 1 : 0);
  if (data_lines % 4 != 0) { *err = "malformed FASTQ: line count is not a multiple of 4"; return false; }
  const uint64_t nrec = data_lines / 4;
  starts->assign(nrec + 1, 0);
  auto emit = [&](size_t c) {
    const uint8_t* base = m.p;
    const uint8_t* p = base + c * chunk;
    const uint8_t* e = base + std::min(m.len, (c + 1) * chunk);
    uint64_t line = nl[c];  // index of the line that contains / starts at p's predecessor newline
    // a line starts at p iff p == base or p[-1] == '\n'
    if (c == 0) (*starts)[0] = 0;
    while (p < e) {
      const void* q = memchr(p, '\n', e - p);
      if (!q) break;
      ++line;  // `line` newlines seen so far: the next line has index `line`
      const uint64_t pos = (const uint8_t*)q - base + 1;
      if (line % 4 == 0 && line / 4 <= nrec) (*starts)[line / 4] = pos;
      p = (const uint8_t*)q + 1;
    }
  };
  {
    std::vector<std::thread> th;
    for (size_t c = 0; c < nchunks; ++c) th.emplace_back(emit, c);
    for (auto& t : th) t.join();
  }
  if ((*starts)[nrec] == 0 && nrec > 0) (*starts)[nrec] = m.len;  // last record without trailing EOL
  // '@' / '+' / equal lengths are checked per batch by bk_tokenize
  return true;
}

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

struct Batch { uint64_t rec0, rec1; };

inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

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
  if (a->compression != 0) {
    set_err("compressed output (--gz/--bgz/--mgz/--lz4) is not supported by this build", err, errcap);
    return BK_E_UNSUPPORTED;
  }
  std::string e;
  Logger log;
  log.quiet = a->quiet != 0;
  log.message("Estimating reads count...");  // run.rs:97 / :215
  const bool dbg = a->verbose != 0;
  const double t_begin = now_s();

  // the CUDA contexts come up while the host scans the input
  std::thread warm([&] {
    const int nd = bk_device_count();
    const int ng = a->n_gpus > 0 ? std::min(a->n_gpus, nd) : std::min(nd, 1);  // (further devices come up in their workers)
    for (int d = 0; d < ng; ++d) bk_warmup(d);
  });
  struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } warm_join{warm};
  Mapped in1, in2;
  if (!map_file(a->fq1, &in1, &e)) { set_err(e, err, errcap); return BK_E_IO; }
  if (pe && !map_file(a->fq2, &in2, &e)) { set_err(e, err, errcap); return BK_E_IO; }
  for (const Mapped* m : {&in1, &in2}) {
    if (const char* k = m->p ? compressed_kind(*m) : nullptr) {
      set_err(std::string(k) + " input is not supported by this build (plain FASTQ only)", err, errcap);
      return BK_E_UNSUPPORTED;
    }
  }
  const unsigned host_threads = std::max<unsigned>(1, std::max<unsigned>((unsigned)a->threads, std::thread::hardware_concurrency()));
  std::vector<uint64_t> st1, st2;
  if (!record_starts(in1, host_threads, &st1, &e)) { set_err(e, err, errcap); return BK_E_FORMAT; }
  if (pe && !record_starts(in2, host_threads, &st2, &e)) { set_err(e, err, errcap); return BK_E_FORMAT; }
  uint64_t nrec = st1.size() - 1;
  if (pe) nrec = std::min<uint64_t>(nrec, st2.size() - 1);  // run.rs:172-173 zip truncates
  if (dbg) fprintf(stderr, "[bk_run] map + record starts: %.3f s\n", now_s() - t_begin);

  int fd1 = open_output(a->out_fq1, a->force != 0, &e);  // run.rs:106-111 / :225-229
  if (fd1 < 0) { set_err(e, err, errcap); return BK_E_IO; }
  int fd2 = -1;
  if (pe) {
    fd2 = open_output(a->out_fq2, a->force != 0, &e);
    if (fd2 < 0) { close(fd1); set_err(e, err, errcap); return BK_E_IO; }
  }
  auto close_outputs = [&]() { close(fd1); if (fd2 >= 0) close(fd2); };

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
  int ndev = bk_device_count();
  if (ndev <= 0) {
    free_programs(); close_outputs();
    set_err("no CUDA device available; this build has no CPU path", err, errcap);
    return BK_E_CUDA;
  }
  int ngpu = a->n_gpus > 0 ? std::min(a->n_gpus, ndev) : ndev;

  // batches: contiguous record ranges, sized by bytes (-m MB per mate when given, else 64 MiB: small enough to pipeline, large enough for the kernel)
  size_t batch_bytes = a->max_memory_mb ? a->max_memory_mb << 20 : (size_t)64 << 20;
  batch_bytes = std::min<size_t>(std::max<size_t>(batch_bytes, 1 << 20), (size_t)3 << 30);
  std::vector<Batch> batches;
  for (uint64_t r = 0; r < nrec;) {
    uint64_t lo = r, hi = nrec;  // largest r1 with both mates' bytes <= batch_bytes
    auto fits = [&](uint64_t r1) {
      return st1[r1] - st1[r] <= batch_bytes && (!pe || st2[r1] - st2[r] <= batch_bytes) && r1 - r < (1u << 30);
    };
    while (lo < hi) {
      uint64_t mid = lo + (hi - lo + 1) / 2;
      if (fits(mid)) lo = mid; else hi = mid - 1;
    }
    if (lo == r) lo = r + 1;  // a single record larger than the budget still goes through
    batches.push_back({r, lo});
    r = lo;
  }
  ngpu = std::max(1, std::min<int>(ngpu, (int)std::max<size_t>(batches.size(), 1)));
  // Unless the caller asked for a number of GPUs: one per 32 batches (2 GiB of text per mate).  The files, not the
  // device, bound this path (one GPU streams ~9 M pairs/s against 3.2 G on the device), and every further
  // context costs its start-up and pinned buffers (2 GPUs on a 2 GB input: 2.2 s instead of 1.5 s).
  if (a->n_gpus <= 0) ngpu = std::max(1, std::min<int>(ngpu, (int)(batches.size() / 32)));

  std::atomic<size_t> next_batch{0};
  std::mutex wmu;
  std::condition_variable wcv;
  size_t next_to_write = 0;
  std::atomic<int> status{BK_OK};
  std::string first_error;
  std::mutex emu;
  auto fail = [&](int code, const std::string& m) {
    {
      std::lock_guard<std::mutex> g(emu);
      if (status.load() == BK_OK) { first_error = m; status = code; }
    }
    // The writers test `status` under wmu before they block: taking it here closes the window between
    // their test and their wait, in which a notification would be lost and the run would hang.
    { std::lock_guard<std::mutex> lk(wmu); }
    wcv.notify_all();
  };
  const uint32_t flags = (pe ? BK_FLAG_PAIRED : 0) | (a->skip_trimming ? BK_FLAG_SKIP_TRIM : 0) |
                         (a->rc_barcodes ? BK_FLAG_RC : 0) | (dbg ? BK_FLAG_VERBOSE : 0) |
                         (a->ctx_flags & (BK_FLAG_DEVICE_BOTH_MATES | BK_FLAG_HOST_MATE_PINNED));

  // One worker per GPU: stage batch k (all host threads: copy into pinned memory, offsets, record checks)
  // -> submit k -> wait k-1 -> hand k-1 to the worker's writer thread, so that staging, the device and the
  // output files all work at the same time.  Three host buffer sets per worker, two device slots (bk_ctx).
  const unsigned stage_threads = std::max(1u, std::min(host_threads / (unsigned)ngpu, 32u));
  auto worker = [&](int dev) {
    char emsg[512];
    int cst = BK_OK;
    double t_ctx = now_s(), t_stage = 0, t_submit = 0, t_wait = 0, t_write = 0, t_setwait = 0;
    bk_ctx* ctx = bk_ctx_create(dev, p1, p2, flags, &cst, emsg, sizeof emsg);
    if (!ctx) { fail(cst, emsg); return; }
    t_ctx = now_s() - t_ctx;
    const int nm = pe ? 2 : 1;
    constexpr int kSets = 3;
    struct HostSet { uint8_t* text[2] = {nullptr, nullptr}; uint32_t* off[2] = {nullptr, nullptr};
                     uint8_t* out[2] = {nullptr, nullptr};
                     size_t tcap[2] = {0, 0}, ocap[2] = {0, 0}, fcap[2] = {0, 0};
                     size_t batch = SIZE_MAX; bk_result res; bool busy = false; } hs[kSets];
    std::mutex smu;                 // guards hs[].busy, the writer's queue
    std::condition_variable scv;
    std::vector<int> wq;            // sets waiting to be written, in submission order
    bool wdone = false;
    double t_alloc = 0;
    auto ensure_host = [&](void** p, size_t* cap, size_t need) {
      if (*cap >= need) return true;
      const double t0 = now_s();
      if (*p) bk_host_free(*p);
      *p = bk_host_alloc(need + need / 8 + 4096);
      *cap = *p ? need + need / 8 + 4096 : 0;
      t_alloc += now_s() - t0;
      return *p != nullptr;
    };
    // the BK_E_FORMAT conditions of bk_tokenize for one record [b, e) (its four lines are given by the record starts)
    auto record_ok = [](const uint8_t* b, const uint8_t* e) {
      if (e - b < 4 || b[0] != '@') return false;
      const uint8_t* nl1 = (const uint8_t*)memchr(b, '\n', e - b);
      if (!nl1 || nl1 + 1 >= e) return false;
      const uint8_t* nl2 = (const uint8_t*)memchr(nl1 + 1, '\n', e - nl1 - 1);
      if (!nl2 || nl2 + 1 >= e || nl2[1] != '+') return false;
      const uint8_t* nl3 = (const uint8_t*)memchr(nl2 + 1, '\n', e - nl2 - 1);
      if (!nl3) return false;
      const uint8_t* nl4 = nl3 + 1 < e ? (const uint8_t*)memchr(nl3 + 1, '\n', e - nl3 - 1) : nullptr;
      if (!nl4) nl4 = e;  // the file's last record may lack its EOL
      size_t sl = nl2 - (nl1 + 1), ql = nl4 - (nl3 + 1);
      if (sl && nl2[-1] == '\r') --sl;
      if (ql && nl4[-1] == '\r') --ql;
      return sl == ql;
    };
    auto stage = [&](int hi, size_t bi, bk_mate_in* in) -> bool {
      HostSet& h = hs[hi];
      const Batch& b = batches[bi];
      const uint32_t n = (uint32_t)(b.rec1 - b.rec0);
      const Mapped* mp[2] = {&in1, &in2};
      const std::vector<uint64_t>* stp[2] = {&st1, &st2};
      for (int m = 0; m < nm; ++m) {
        const uint64_t t0 = (*stp[m])[b.rec0], t1 = (*stp[m])[b.rec1];
        if (!ensure_host((void**)&h.text[m], &h.tcap[m], t1 - t0 + 64)) return false;
        if (!ensure_host((void**)&h.off[m], &h.fcap[m], ((size_t)n + 2) * 4)) return false;
        if (!ensure_host((void**)&h.out[m], &h.ocap[m], bk_out_bound(ctx, m + 1, n, t1 - t0))) return false;
      }
      const unsigned T = n < 20000 ? 1u : stage_threads;
      std::vector<uint32_t> maxrec((size_t)T * 2, 0);
      std::atomic<bool> bad{false};
      auto part = [&](unsigned t) {
        const uint64_t r0 = b.rec0 + (uint64_t)n * t / T, r1 = b.rec0 + (uint64_t)n * (t + 1) / T;
        for (int m = 0; m < nm; ++m) {
          const std::vector<uint64_t>& st = *stp[m];
          const uint64_t base = st[b.rec0];
          memcpy(h.text[m] + (st[r0] - base), mp[m]->p + st[r0], st[r1] - st[r0]);
          uint32_t mr = 0;
          for (uint64_t r = r0; r < r1; ++r) {
            h.off[m][r - b.rec0] = (uint32_t)(st[r] - base);
            const uint64_t len = st[r + 1] - st[r];
            if (len > mr) mr = (uint32_t)len;
            if (!record_ok(mp[m]->p + st[r], mp[m]->p + st[r + 1])) bad = true;
          }
          if (t + 1 == T) h.off[m][n] = (uint32_t)(st[b.rec1] - base);
          maxrec[(size_t)t * 2 + m] = mr;
        }
      };
      if (T == 1) part(0);
      else {
        std::vector<std::thread> th;
        for (unsigned t = 1; t < T; ++t) th.emplace_back(part, t);
        part(0);
        for (auto& x : th) x.join();
      }
      if (bad.load()) { fail(BK_E_FORMAT, "malformed FASTQ record"); return false; }
      for (int m = 0; m < nm; ++m) {
        uint32_t mr = 1;
        for (unsigned t = 0; t < T; ++t) mr = std::max(mr, maxrec[(size_t)t * 2 + m]);
        in[m].text = h.text[m];
        in[m].rec_off = h.off[m];
        in[m].text_len = (*stp[m])[b.rec1] - (*stp[m])[b.rec0];
        in[m].max_rec_bytes = mr;
        in[m].reserved = 0;
      }
      h.batch = bi;
      return true;
    };
    // writer: sets in submission order; the files in batch order across all workers (run.rs:79,194)
    std::thread writer([&] {
      for (;;) {
        int hi;
        {
          std::unique_lock<std::mutex> lk(smu);
          scv.wait(lk, [&] { return !wq.empty() || wdone; });
          if (wq.empty()) return;
          hi = wq.front();
          wq.erase(wq.begin());
        }
        HostSet& h = hs[hi];
        const double tw = now_s();
        bool ok = true;
        {
          std::unique_lock<std::mutex> lk(wmu);
          wcv.wait(lk, [&] { return next_to_write == h.batch || status.load() != BK_OK; });
          ok = status.load() == BK_OK;
        }
        if (ok) {  // (only the holder of the turn writes: no lock needed while it does)
          std::thread w2;
          bool ok2 = true;
          if (pe) w2 = std::thread([&] { ok2 = write_all(fd2, h.out[1], h.res.out2_len); });
          ok = write_all(fd1, h.out[0], h.res.out1_len);
          if (pe) w2.join();
          ok = ok && ok2;
          if (!ok) fail(BK_E_IO, io_error(strerror(errno)));
          {
            std::lock_guard<std::mutex> lk(wmu);
            ++next_to_write;
          }
          wcv.notify_all();
        }
        t_write += now_s() - tw;
        {
          std::lock_guard<std::mutex> lk(smu);
          h.busy = false;
          h.batch = SIZE_MAX;
        }
        scv.notify_all();
      }
    });
    auto finish = [&](int dslot, int hi) -> bool {  // wait for the device, hand the set to the writer
      HostSet& h = hs[hi];
      const double tw0 = now_s();
      int rc = bk_wait(ctx, dslot, h.out[0], h.ocap[0], h.out[1], h.ocap[1], &h.res);
      t_wait += now_s() - tw0;
      if (rc != BK_OK) { fail(rc, bk_last_error(ctx)); return false; }
      {
        std::lock_guard<std::mutex> lk(smu);
        wq.push_back(hi);
      }
      scv.notify_all();
      return true;
    };
    bool ok = true;
    long k = 0;            // batches this worker has submitted
    int prev_set = -1;     // set of batch k-1, still on the device
    for (;; ++k) {
      size_t bi = next_batch.fetch_add(1);
      if (bi >= batches.size() || status.load() != BK_OK) break;
      const int hi = (int)(k % kSets);
      {
        const double t0 = now_s();
        std::unique_lock<std::mutex> lk(smu);
        scv.wait(lk, [&] { return !hs[hi].busy; });
        hs[hi].busy = true;
        t_setwait += now_s() - t0;
      }
      bk_mate_in in[2];
      const double ts0 = now_s();
      if (!(ok = stage(hi, bi, in))) { if (status.load() == BK_OK) fail(BK_E_CAPACITY, "pinned host allocation failed"); break; }
      const double ts1 = now_s();
      t_stage += ts1 - ts0;
      int rc = bk_submit(ctx, (int)(k & 1), (uint32_t)(batches[bi].rec1 - batches[bi].rec0), &in[0], pe ? &in[1] : nullptr);
      t_submit += now_s() - ts1;
      if (rc != BK_OK) { fail(rc, bk_last_error(ctx)); ok = false; break; }
      if (prev_set >= 0 && !(ok = finish((int)((k - 1) & 1), prev_set))) { prev_set = -1; break; }
      prev_set = hi;
    }
    if (ok && prev_set >= 0) ok = finish((int)((k - 1) & 1), prev_set);
    {
      std::lock_guard<std::mutex> lk(smu);
      wdone = true;
    }
    scv.notify_all();
    writer.join();
    for (auto& h : hs)
      for (int m = 0; m < 2; ++m) {
        if (h.text[m]) bk_host_free(h.text[m]);
        if (h.off[m]) bk_host_free(h.off[m]);
        if (h.out[m]) bk_host_free(h.out[m]);
      }
    bk_ctx_destroy(ctx);
    if (dbg) fprintf(stderr, "[bk_run] gpu %d: context %.3f s, stage %.3f s (%u threads), submit %.3f s, wait %.3f s, "
                     "writer busy %.3f s, waiting for a free buffer set %.3f s; pinned allocation inside stage %.3f s\n", dev, t_ctx, t_stage,
                     stage_threads, t_submit, t_wait, t_write, t_setwait, t_alloc);
  };

  if (!batches.empty()) {
    std::vector<std::thread> th;
    for (int d = 0; d < ngpu; ++d) th.emplace_back(worker, d);
    for (auto& t : th) t.join();
  }
  free_programs();
  close_outputs();
  if (dbg) fprintf(stderr, "[bk_run] total %.3f s, %zu batches\n", now_s() - t_begin, batches.size());
  if (status.load() != BK_OK) { set_err(first_error, err, errcap); return status.load(); }
  log.final_message();
  return BK_OK;
}
