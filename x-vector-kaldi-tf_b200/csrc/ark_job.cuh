// ark_job.cuh -- host side of an ark -> x-vector-ark extraction job (include/xvec_job.h): striped header index,
// make_embedding's skip / chunk rules, a pread pool that fills page-locked batches, and the vector-ark / scp formatters.
// Host-only; included at the end of xvec_api.cu (shares fail() / XV_CUDA and xv_ark_scan).
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>

#include "../../include/xvec_job.h"
#include "synth.cuh"

extern "C" void xv_convert_f32_to_f16_host(const float* src, uint16_t* dst, size_t n);     // f16_convert.cpp

namespace arkjob {

struct Entry {
  int64_t key_off;     // in the file
  int64_t key_pos;     // in the reader's key blob (the scan keeps the key bytes it has read anyway)
  int64_t payload_off;
  int32_t key_len;
  int32_t rows;
  int32_t elem;      // 4 = float32, 8 = float64
  int32_t used;      // rows kept (chunks are contiguous from row 0; only a tail shorter than min_chunk_size is dropped)
  int32_t n_chunks;  // 0: skipped
  int32_t reason;    // XV_UTT_*
};

struct Batch {
  int64_t u0, u1;        // ok-utterance range
  int64_t seg0;          // first segment (global over the stripe)
  int64_t first0;        // offset of this batch's utt_first_seg run in first_seg_all (n_utt + 1 entries)
  int64_t n_rows;
  int32_t n_seg;
  int32_t n_tasks;
};

struct Task { int32_t batch; int64_t u0, u1; };

constexpr int64_t TASK_BYTES = 2 << 20;       // payload bytes one worker fetches per claim
constexpr int VALIDATE_ENTRIES = 4;           // headers that must chain behind a resynchronisation candidate

}  // namespace arkjob

struct xv_ark_reader {
  xv_ark_reader_opts o{};
  int fd = -1;
  int64_t file_size = 0;
  const uint8_t* map = nullptr;
  std::vector<arkjob::Entry> entries;          // every matrix of the stripe, in file order
  std::string key_blob;                        // their keys (Entry.key_pos, key_len)
  std::vector<int64_t> ok;                     // entry index of the i-th ok utterance
  std::vector<int64_t> row_in_batch;           // [n_ok] first row of the utterance inside its batch buffer
  std::vector<int32_t> seg_len_all;            // [n_segments]
  std::vector<int32_t> first_seg_all;          // batch-relative utt_first_seg runs, (n_utt + 1) per batch
  std::vector<int64_t> dst_row_all;            // [n_ok]
  std::vector<arkjob::Batch> batches;
  std::vector<arkjob::Task> tasks;
  xv_ark_index_info info{};
  int64_t scan_from = 0;                       // where the stripe's scan starts (a boundary or a candidate's key)
  bool first_is_candidate = false;             // entries[0] came from a resynchronisation (its key start is provisional)
  // ring
  std::vector<float*> slot_buf;
  std::vector<size_t> slot_bytes;
  int64_t slot_rows = 0;
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_work, cv_ready;
  size_t next_task = 0;
  std::vector<int32_t> batch_done;             // tasks finished per batch
  int64_t released = 0;                        // batches handed back by the consumer (in order)
  int64_t next_batch = 0;                      // next batch the consumer gets
  bool stop = false, started = false;
  std::string error;
};

namespace arkjob {

// Page-locked batch buffers are expensive to create (cudaHostAlloc pins every page: ~10 ms per 40 MB) and a process that
// runs several jobs needs the same sizes again: closed readers park their buffers here (at most POOL_CAP_BYTES).
struct PinnedPool {
  std::mutex mu;
  std::vector<std::pair<float*, size_t>> free_list;
  size_t bytes = 0;
  static constexpr size_t POOL_CAP_BYTES = size_t(1) << 30;
  float* take(size_t need) {
    std::lock_guard<std::mutex> lk(mu);
    int best = -1;
    for (int i = 0; i < int(free_list.size()); ++i)
      if (free_list[i].second >= need && (best < 0 || free_list[i].second < free_list[best].second)) best = i;
    if (best < 0 || free_list[best].second > 2 * need + (size_t(1) << 20)) return nullptr;
    float* p = free_list[best].first;
    bytes -= free_list[best].second;
    free_list.erase(free_list.begin() + best);
    return p;
  }
  bool give(float* p, size_t n) {
    std::lock_guard<std::mutex> lk(mu);
    if (bytes + n > POOL_CAP_BYTES) return false;
    free_list.emplace_back(p, n);
    bytes += n;
    return true;
  }
};
inline PinnedPool& pinned_pool() { static PinnedPool pool; return pool; }

inline bool key_char(uint8_t c) { return ark_key_char(c); }

// A "\0B[FD]M \4 rows \4 cols" header at h whose matrix has the expected width and fits in the file.
bool plausible_header(const xv_ark_reader* r, int64_t h, int64_t* payload_end) {
  const uint8_t* b = r->map;
  if (h < 2 || h + 15 > r->file_size) return false;
  if (b[h] != 0 || b[h + 1] != 'B' || (b[h + 2] != 'F' && b[h + 2] != 'D') || b[h + 3] != 'M' || b[h + 4] != ' ' ||
      b[h + 5] != 4 || b[h + 10] != 4 || b[h - 1] != ' ' || !key_char(b[h - 2]))
    return false;
  int32_t rows, cols;
  memcpy(&rows, b + h + 6, 4);
  memcpy(&cols, b + h + 11, 4);
  if (rows < 0 || cols != r->o.feat_dim) return false;
  const int64_t end = h + 15 + int64_t(rows) * cols * (b[h + 2] == 'F' ? 4 : 8);
  if (end > r->file_size) return false;
  *payload_end = end;
  return true;
}

// First marker offset >= from whose header is plausible AND behind which VALIDATE_ENTRIES further entries (or the end of
// the file) parse in a chain.  -1: none.
int64_t resync(const xv_ark_reader* r, int64_t from) {
  const uint8_t* b = r->map;
  int64_t pos = std::max<int64_t>(from, 2);
  while (pos + 15 <= r->file_size) {
    const void* z = memchr(b + pos, 0, size_t(r->file_size - pos));
    if (!z) return -1;
    const int64_t h = static_cast<const uint8_t*>(z) - b;
    int64_t end = 0;
    if (plausible_header(r, h, &end)) {
      int64_t ko[VALIDATE_ENTRIES], po[VALIDATE_ENTRIES], consumed = 0;
      int32_t kl[VALIDATE_ENTRIES], rw[VALIDATE_ENTRIES], cl[VALIDATE_ENTRIES], eb[VALIDATE_ENTRIES];
      const int64_t n = xv_ark_scan(b + end, r->file_size - end, VALIDATE_ENTRIES, ko, kl, rw, cl, eb, po, &consumed);
      bool ok = n == VALIDATE_ENTRIES || (n >= 0 && end + consumed == r->file_size);
      for (int64_t i = 0; i < n && ok; ++i) ok = cl[i] == r->o.feat_dim;
      if (ok) return h;
    }
    pos = h + 1;
  }
  return -1;
}

// make_embedding's rules for one utterance (models.py:377-409).
void plan_entry(const xv_ark_reader_opts& o, Entry* e) {
  const int64_t rows = e->rows;
  e->used = 0; e->n_chunks = 0; e->reason = XV_UTT_OK;
  if (rows == 0) { e->reason = XV_UTT_ZERO_LENGTH; return; }
  if (rows < o.min_chunk_size) { e->reason = XV_UTT_TOO_SHORT; return; }
  if (o.chunk_size == -1 || rows <= o.chunk_size) { e->used = int32_t(rows); e->n_chunks = 1; return; }
  const int64_t cs = o.chunk_size;
  const int64_t n = (rows + cs - 1) / cs;
  const int64_t tail = rows - (n - 1) * cs;
  if (tail < o.min_chunk_size) { e->n_chunks = int32_t(n - 1); e->used = int32_t((n - 1) * cs); }
  else { e->n_chunks = int32_t(n); e->used = int32_t(rows); }
}

int build_plan(xv_ark_reader* r) {
  const xv_ark_reader_opts& o = r->o;
  r->ok.clear(); r->row_in_batch.clear(); r->seg_len_all.clear(); r->first_seg_all.clear(); r->batches.clear(); r->tasks.clear();
  xv_ark_index_info& inf = r->info;
  inf.n_entries = int64_t(r->entries.size());
  inf.n_ok = inf.n_fail = inf.n_segments = inf.rows_used = inf.key_bytes = 0;
  Batch cur{};
  bool open = false;
  auto close_batch = [&]() {
    if (!open) return;
    cur.u1 = int64_t(r->ok.size());
    r->first_seg_all.push_back(cur.n_seg);
    // tasks: runs of utterances of about TASK_BYTES
    int64_t bytes = 0, t0 = cur.u0;
    cur.n_tasks = 0;
    for (int64_t u = cur.u0; u < cur.u1; ++u) {
      bytes += int64_t(r->entries[r->ok[u]].used) * o.feat_dim * 4;
      if (bytes >= TASK_BYTES || u + 1 == cur.u1) {
        r->tasks.push_back(Task{int32_t(r->batches.size()), t0, u + 1});
        ++cur.n_tasks;
        t0 = u + 1;
        bytes = 0;
      }
    }
    r->batches.push_back(cur);
    open = false;
  };
  for (size_t i = 0; i < r->entries.size(); ++i) {
    Entry& e = r->entries[i];
    plan_entry(o, &e);
    if (e.n_chunks == 0) { ++inf.n_fail; continue; }
    // the first batches are small, so that the device starts after a fraction of a millisecond of reading: 1/8, 1/4, 1/2
    // of batch_frames, then full batches (a batch of 50 000 frames already runs the kernels at full efficiency)
    const int64_t nb = int64_t(r->batches.size());
    const int64_t target = (o.batch_frames >= 131072 && nb < 3) ? (o.batch_frames >> (3 - nb)) : o.batch_frames;
    if (open && cur.n_rows + e.used > std::max<int64_t>(target, e.used)) close_batch();
    if (!open) {
      cur = Batch{};
      cur.u0 = int64_t(r->ok.size());
      cur.seg0 = int64_t(r->seg_len_all.size());
      cur.first0 = int64_t(r->first_seg_all.size());
      open = true;
    }
    r->row_in_batch.push_back(cur.n_rows);
    r->first_seg_all.push_back(cur.n_seg);
    const int64_t cs = o.chunk_size;
    for (int32_t c = 0; c < e.n_chunks; ++c) {
      const int32_t len = e.n_chunks == 1 && (cs == -1 || e.rows <= cs) ? e.used : int32_t(std::min<int64_t>(cs, e.rows - int64_t(c) * cs));
      r->seg_len_all.push_back(len);
    }
    cur.n_seg += e.n_chunks;
    cur.n_rows += e.used;
    r->ok.push_back(int64_t(i));
    ++inf.n_ok;
    inf.n_segments += e.n_chunks;
    inf.rows_used += e.used;
    inf.key_bytes += e.key_len;
  }
  close_batch();
  inf.n_batches = int64_t(r->batches.size());
  inf.first_marker_off = r->entries.empty() ? -1 : r->entries[0].payload_off - 15;
  return XV_OK;
}

// Chain scan of [from, ...): the entries whose marker lies below end_limit, and what follows them.  Thread-safe (no fail()).
struct RangeScan {
  std::vector<Entry> entries;
  std::string keys;
  int64_t next_marker = 0, next_key = 0, stopped_at = -1;
  std::string err;
};

// The chain is walked with one small pread per header (a few hundred bytes: key + 15 header bytes) rather than through the
// mapping: a page fault per header (~1 us, more on some hosts: 8-10 ms per 10 000 utterances were measured) costs more than
// the system call.
void scan_range(const xv_ark_reader* r, int64_t from, int64_t end_limit, RangeScan* out) {
  const xv_ark_reader_opts& o = r->o;
  out->entries.clear();
  out->keys.clear();
  out->stopped_at = -1;
  out->next_marker = out->next_key = r->file_size;
  out->err.clear();
  uint8_t small[320];
  std::vector<uint8_t> big;
  int64_t pos = from;
  while (pos < r->file_size) {
    const uint8_t* buf = small;
    int64_t want = std::min<int64_t>(int64_t(sizeof(small)), r->file_size - pos), avail = 0;
    int64_t ko = 0, po = 0;
    int32_t kl = 0, rw = 0, cl = 0, eb = 0;
    int st;
    for (;;) {
      uint8_t* dst = buf == small ? small : big.data();
      while (avail < want) {
        const ssize_t k = pread(r->fd, dst + avail, size_t(want - avail), off_t(pos + avail));
        if (k <= 0) break;
        avail += k;
      }
      st = ark_parse_header(dst, avail, r->file_size - pos, &ko, &kl, &rw, &cl, &eb, &po);
      if (st != -1 || avail < want || buf != small) break;
      big.resize(4300);                                  // a very long key: one more, larger read
      memcpy(big.data(), small, size_t(avail));
      buf = big.data();
      want = std::min<int64_t>(4300, r->file_size - pos);
    }
    if (st != 1) {
      // trailing white space is tolerated (a text-mode tail); anything else is an entry this scanner does not know
      const uint8_t* p = buf == small ? small : big.data();
      int64_t q = 0;
      while (q < avail && ark_space(p[q])) ++q;
      if (!(q == avail && pos + avail == r->file_size)) out->stopped_at = pos;
      return;
    }
    const int64_t marker = pos + po - 15;
    if (marker >= end_limit) {
      out->next_marker = marker;
      out->next_key = pos + ko;
      return;
    }
    if (cl != o.feat_dim && rw > 0) {
      const uint8_t* p = buf == small ? small : big.data();
      out->err = "utterance " + std::string(reinterpret_cast<const char*>(p + ko), size_t(kl)) + " has feature dim " +
                 std::to_string(cl) + ", model expects " + std::to_string(o.feat_dim);
      return;
    }
    Entry e{};
    e.key_off = pos + ko; e.key_len = kl; e.rows = rw; e.elem = eb; e.payload_off = pos + po;
    e.key_pos = int64_t(out->keys.size());
    out->keys.append(reinterpret_cast<const char*>((buf == small ? small : big.data()) + ko), size_t(kl));
    out->entries.push_back(e);
    pos = e.payload_off + int64_t(rw) * cl * eb;
  }
}

// The first entry of a resynchronised range was parsed from a provisional key start; the chain in front knows the true one.
void fix_first_key(const xv_ark_reader* r, Entry* e, std::string* keys, int64_t marker, int64_t key_off) {
  const int64_t len = marker - 1 - key_off;
  e->key_off = key_off;
  e->key_len = int32_t(len);
  e->key_pos = int64_t(keys->size());
  keys->append(reinterpret_cast<const char*>(r->map + key_off), size_t(len));
}

// Provisional start of the key in front of the marker at h: the longest run of key characters before the separating space
// (the previous payload's last bytes may look like key characters too; the chain of the range in front knows the truth).
int64_t provisional_key(const xv_ark_reader* r, int64_t h) {
  int64_t k = h - 1;
  while (k > 0 && key_char(r->map[k - 1]) && h - k < 4096) --k;
  return k;
}

// Index of the stripe [from, end_limit).  The chain is sequential by nature (a header tells where the next one is) and
// every header costs a page fault of the mapping (~1 us: 10 ms per 10 000 utterances), so the stripe is cut into
// sub-ranges that are scanned in parallel from resynchronised candidates and then stitched: a sub-range whose first
// marker is not the one its predecessor's chain arrives at is scanned again from the true boundary.
int scan_stripe(xv_ark_reader* r, int64_t from, bool first_is_candidate) {
  const xv_ark_reader_opts& o = r->o;
  const int64_t end_limit = o.byte_end < 0 ? r->file_size : std::min<int64_t>(o.byte_end, r->file_size);
  r->entries.clear();
  r->scan_from = from;
  r->first_is_candidate = first_is_candidate;
  xv_ark_index_info& inf = r->info;
  inf = xv_ark_index_info{};
  inf.stopped_at = -1;
  constexpr int64_t MIN_PART = int64_t(4) << 20;
  const int parts = int(std::max<int64_t>(1, std::min<int64_t>(std::min(o.n_threads, 16), (end_limit - from) / MIN_PART)));
  std::vector<RangeScan> scans(size_t(parts), RangeScan{});
  std::vector<int64_t> part_end(size_t(parts), end_limit), first_marker(size_t(parts), -1);
  for (int i = 0; i + 1 < parts; ++i) part_end[i] = from + (end_limit - from) * (i + 1) / parts;
  auto run_part = [&](int i) {
    int64_t start = from;
    if (i > 0) {
      const int64_t h = resync(r, part_end[i - 1]);
      if (h < 0 || h >= part_end[i]) { scans[i].next_marker = scans[i].next_key = -1; return; }   // nothing found in the sub-range
      first_marker[i] = h;
      start = provisional_key(r, h);
    }
    scan_range(r, start, part_end[i], &scans[i]);
  };
  if (parts == 1) run_part(0);
  else {
    std::vector<std::thread> th;
    for (int i = 1; i < parts; ++i) th.emplace_back(run_part, i);
    run_part(0);
    for (auto& t : th) t.join();
  }
  // stitch
  RangeScan& head = scans[0];
  if (!head.err.empty()) return fail(XV_EINVAL, head.err);
  r->entries.swap(head.entries);
  r->key_blob.swap(head.keys);
  int64_t next_marker = head.next_marker, next_key = head.next_key, stopped = head.stopped_at;
  for (int i = 1; i < parts && stopped < 0; ++i) {
    if (next_marker >= part_end[i]) continue;                  // the chain in front runs past this whole sub-range
    RangeScan& s = scans[i];
    if (first_marker[i] != next_marker || s.entries.empty()) {
      scan_range(r, next_key, part_end[i], &s);                // the candidate was wrong (or missing): true boundary
    } else {
      fix_first_key(r, &s.entries[0], &s.keys, next_marker, next_key);
    }
    if (!s.err.empty()) return fail(XV_EINVAL, s.err);
    const int64_t shift = int64_t(r->key_blob.size());
    r->key_blob += s.keys;
    for (auto& e : s.entries) e.key_pos += shift;
    r->entries.insert(r->entries.end(), s.entries.begin(), s.entries.end());
    next_marker = s.next_marker; next_key = s.next_key; stopped = s.stopped_at;
  }
  inf.next_marker_off = next_marker;
  inf.next_key_off = next_key;
  inf.stopped_at = stopped;
  return build_plan(r);
}

void worker_main(xv_ark_reader* r) {
  std::vector<double> tmp;
  std::vector<float> ftmp;
  for (;;) {
    Task t;
    {
      std::unique_lock<std::mutex> lk(r->mu);
      for (;;) {
        if (r->stop || !r->error.empty()) return;
        if (r->next_task < r->tasks.size() && r->tasks[r->next_task].batch < r->released + int64_t(r->slot_buf.size())) break;
        if (r->next_task >= r->tasks.size()) return;
        r->cv_work.wait(lk);
      }
      t = r->tasks[r->next_task++];
    }
    const Batch& b = r->batches[t.batch];
    float* base = r->slot_buf[t.batch % r->slot_buf.size()];
    const bool f16 = r->o.feats_f16 != 0;
    std::string err;
    for (int64_t u = t.u0; u < t.u1 && err.empty(); ++u) {
      const Entry& e = r->entries[r->ok[u]];
      const int64_t n_val = int64_t(e.used) * r->o.feat_dim;
      const int64_t first = r->row_in_batch[u] * r->o.feat_dim;
      auto read_all = [&](uint8_t* p, int64_t need, int64_t file_off) {
        int64_t got = 0;
        while (got < need) {
          const ssize_t k = pread(r->fd, p + got, size_t(need - got), off_t(file_off + got));
          if (k <= 0) { err = "truncated matrix payload at byte " + std::to_string(file_off + got); return false; }
          got += k;
        }
        return true;
      };
      if (e.elem == 4 && !f16) {                       // float32 payload -> float32 batch: straight into the page-locked buffer
        read_all(reinterpret_cast<uint8_t*>(base + first), n_val * 4, e.payload_off);
      } else {
        // through a cache-resident bounce buffer: narrow float64 payloads and / or round to float16 on the way
        constexpr int64_t CHUNK = 16384;               // values per step (64 / 128 KB of payload)
        if (int64_t(tmp.size()) < CHUNK) { tmp.resize(size_t(CHUNK)); ftmp.resize(size_t(CHUNK)); }
        uint16_t* hbase = reinterpret_cast<uint16_t*>(base);
        for (int64_t v0 = 0; v0 < n_val && err.empty(); v0 += CHUNK) {
          const int64_t n = std::min(CHUNK, n_val - v0);
          const float* src;
          if (e.elem == 8) {
            if (!read_all(reinterpret_cast<uint8_t*>(tmp.data()), n * 8, e.payload_off + v0 * 8)) break;
            for (int64_t i = 0; i < n; ++i) ftmp[size_t(i)] = float(tmp[size_t(i)]);
            src = ftmp.data();
          } else {
            if (!read_all(reinterpret_cast<uint8_t*>(ftmp.data()), n * 4, e.payload_off + v0 * 4)) break;
            src = ftmp.data();
          }
          if (f16) xv_convert_f32_to_f16_host(src, hbase + first + v0, size_t(n));
          else memcpy(base + first + v0, src, size_t(n) * 4);
        }
      }
    }
    (void)b;
    {
      std::lock_guard<std::mutex> lk(r->mu);
      if (!err.empty() && r->error.empty()) r->error = err;
      if (++r->batch_done[t.batch] == r->batches[t.batch].n_tasks || !r->error.empty()) r->cv_ready.notify_all();
      if (!r->error.empty()) r->cv_work.notify_all();
    }
  }
}

void stop_workers(xv_ark_reader* r) {
  {
    std::lock_guard<std::mutex> lk(r->mu);
    r->stop = true;
  }
  r->cv_work.notify_all();
  r->cv_ready.notify_all();
  for (auto& w : r->workers) if (w.joinable()) w.join();
  r->workers.clear();
}

}  // namespace arkjob

extern "C" {

int xv_ark_reader_open(xv_ark_reader** out, const char* path, const xv_ark_reader_opts* opts) {
  if (!out || !path || !opts) return fail(XV_EINVAL, "null argument");
  *out = nullptr;
  if (opts->feat_dim <= 0 || opts->n_threads < 1 || opts->n_slots < 2 || opts->batch_frames < 1 || opts->byte_begin < 0 ||
      opts->min_chunk_size < 0 || (opts->chunk_size < 1 && opts->chunk_size != -1))
    return fail(XV_EINVAL, "xv_ark_reader_open: bad options");
  if (opts->chunk_size != -1 && opts->chunk_size < opts->min_chunk_size)
    return fail(XV_EINVAL, "chunk_size is smaller than min_chunk_size: no chunk would ever be evaluated");
  const int fd = open(path, O_RDONLY | O_CLOEXEC);
  if (fd < 0) return fail(XV_EINVAL, std::string("cannot open ") + path + ": " + strerror(errno));
  struct stat st;
  if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) {
    close(fd);
    return fail(XV_EINVAL, std::string(path) + " is not a regular file");
  }
  xv_ark_reader* r = new xv_ark_reader();
  r->o = *opts;
  r->fd = fd;
  r->file_size = int64_t(st.st_size);
  if (r->file_size > 0) {
    void* m = mmap(nullptr, size_t(r->file_size), PROT_READ, MAP_SHARED, fd, 0);
    if (m == MAP_FAILED) {
      close(fd);
      delete r;
      return fail(XV_ENOMEM, std::string("mmap of ") + path + " failed: " + strerror(errno));
    }
    r->map = static_cast<const uint8_t*>(m);
  }
  *out = r;
  return XV_OK;
}

int xv_ark_reader_index(xv_ark_reader* r, xv_ark_index_info* info) {
  if (!r || !info) return fail(XV_EINVAL, "null argument");
  if (r->started) return fail(XV_ESTATE, "the reader has been started");
  int64_t from = std::min(r->o.byte_begin, r->file_size);
  bool candidate = false;
  if (!r->o.begin_is_boundary && from > 0) {
    const int64_t h = arkjob::resync(r, from);
    if (h < 0) from = r->file_size;
    else {
      // provisional key start: the longest run of key characters in front of the separating space (the previous payload's
      // last bytes may look like key characters too; xv_ark_reader_set_first fixes it)
      from = arkjob::provisional_key(r, h);
      candidate = true;
    }
  }
  int rc = arkjob::scan_stripe(r, from, candidate);
  if (rc != XV_OK) return rc;
  *info = r->info;
  return XV_OK;
}

int xv_ark_reader_set_first(xv_ark_reader* r, int64_t marker_off, int64_t key_off, xv_ark_index_info* info) {
  if (!r || !info) return fail(XV_EINVAL, "null argument");
  if (r->started) return fail(XV_ESTATE, "the reader has been started");
  const int64_t end_limit = r->o.byte_end < 0 ? r->file_size : std::min<int64_t>(r->o.byte_end, r->file_size);
  if (marker_off < 0 || key_off < 0 || key_off > marker_off) return fail(XV_EINVAL, "bad boundary");
  if (marker_off >= end_limit) {
    // the previous stripe's chain runs past this stripe: it is empty, and hands the same boundary on
    r->entries.clear();
    r->key_blob.clear();
    r->first_is_candidate = false;
    int rc = arkjob::build_plan(r);
    if (rc != XV_OK) return rc;
    r->info.stopped_at = -1;
    r->info.next_marker_off = marker_off;
    r->info.next_key_off = key_off;
  } else if (!r->entries.empty() && r->entries[0].payload_off - 15 == marker_off) {
    // the candidate was right: fix where its key starts
    arkjob::fix_first_key(r, &r->entries[0], &r->key_blob, marker_off, key_off);
    r->first_is_candidate = false;
    int rc = arkjob::build_plan(r);
    if (rc != XV_OK) return rc;
  } else {
    int rc = arkjob::scan_stripe(r, key_off, false);       // re-index from the true boundary
    if (rc != XV_OK) return rc;
  }
  *info = r->info;
  return XV_OK;
}

int xv_ark_reader_keys(const xv_ark_reader* r, char* blob, int64_t blob_cap, int64_t* key_off) {
  if (!r || !key_off || (!blob && blob_cap > 0)) return fail(XV_EINVAL, "null argument");
  if (blob_cap < r->info.key_bytes) return fail(XV_ENOMEM, "key buffer too small");
  int64_t pos = 0;
  for (size_t i = 0; i < r->ok.size(); ++i) {
    const arkjob::Entry& e = r->entries[r->ok[i]];
    key_off[i] = pos;
    memcpy(blob + pos, r->key_blob.data() + e.key_pos, size_t(e.key_len));
    pos += e.key_len;
  }
  key_off[r->ok.size()] = pos;
  return XV_OK;
}

int xv_ark_reader_failures(const xv_ark_reader* r, int32_t* reason, int32_t* rows, char* blob, int64_t blob_cap, int64_t* key_off) {
  if (!r || !reason || !rows || !key_off) return fail(XV_EINVAL, "null argument");
  int64_t pos = 0, n = 0;
  for (const arkjob::Entry& e : r->entries) {
    if (e.n_chunks != 0) continue;
    if (pos + e.key_len > blob_cap) return fail(XV_ENOMEM, "key buffer too small");
    reason[n] = e.reason; rows[n] = e.rows; key_off[n] = pos;
    memcpy(blob + pos, r->key_blob.data() + e.key_pos, size_t(e.key_len));
    pos += e.key_len;
    ++n;
  }
  key_off[n] = pos;
  return XV_OK;
}

int xv_ark_reader_start(xv_ark_reader* r, int64_t dst_row_base) {
  if (!r) return fail(XV_EINVAL, "null argument");
  if (r->started) return fail(XV_ESTATE, "the reader has been started");
  if (r->first_is_candidate) return fail(XV_ESTATE, "the stripe's first entry is unconfirmed: call xv_ark_reader_set_first");
  r->dst_row_all.resize(r->ok.size());
  for (size_t i = 0; i < r->ok.size(); ++i) r->dst_row_all[i] = dst_row_base + int64_t(i);
  int64_t rows = 1;
  for (const auto& b : r->batches) rows = std::max(rows, b.n_rows);
  r->slot_rows = rows;
  const int n_slots = int(std::min<int64_t>(r->o.n_slots, std::max<int64_t>(int64_t(r->batches.size()), 1)));
  r->slot_buf.assign(size_t(n_slots), nullptr);
  r->slot_bytes.assign(size_t(n_slots), 0);
  for (int s = 0; s < n_slots; ++s) {
    const size_t bytes = size_t(rows) * r->o.feat_dim * (r->o.feats_f16 ? 2 : 4);
    r->slot_bytes[s] = bytes;
    if (r->o.pinned) {
      r->slot_buf[s] = arkjob::pinned_pool().take(bytes);
      if (r->slot_buf[s]) continue;
      cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&r->slot_buf[s]), bytes, cudaHostAllocDefault);
      if (e != cudaSuccess) return fail(XV_ECUDA, std::string("cudaHostAlloc of a batch buffer: ") + cudaGetErrorString(e));
    } else {
      r->slot_buf[s] = static_cast<float*>(malloc(bytes));
      if (!r->slot_buf[s]) return fail(XV_ENOMEM, "out of memory for a batch buffer");
    }
  }
  r->batch_done.assign(r->batches.size(), 0);
  r->started = true;
  const int n_threads = int(std::min<size_t>(size_t(r->o.n_threads), std::max<size_t>(r->tasks.size(), 1)));
  for (int i = 0; i < n_threads; ++i) r->workers.emplace_back(arkjob::worker_main, r);
  return XV_OK;
}

int xv_ark_reader_next(xv_ark_reader* r, xv_ark_batch* batch) {
  if (!r || !batch) return fail(XV_EINVAL, "null argument");
  if (!r->started) return fail(XV_ESTATE, "the reader has not been started");
  *batch = xv_ark_batch{};
  if (r->next_batch >= int64_t(r->batches.size())) return XV_OK;      // n_utt == 0: end of the stripe
  const int64_t bi = r->next_batch;
  {
    std::unique_lock<std::mutex> lk(r->mu);
    r->cv_ready.wait(lk, [&] { return !r->error.empty() || r->batch_done[bi] == r->batches[bi].n_tasks; });
    if (!r->error.empty()) return fail(XV_EINVAL, r->error);
  }
  const arkjob::Batch& b = r->batches[bi];
  batch->slot = int32_t(bi % int64_t(r->slot_buf.size()));
  batch->n_seg = b.n_seg;
  batch->n_utt = int32_t(b.u1 - b.u0);
  batch->feats_f16 = r->o.feats_f16 ? 1 : 0;
  batch->n_rows = b.n_rows;
  batch->feats = r->slot_buf[batch->slot];
  batch->seg_len = r->seg_len_all.data() + b.seg0;
  batch->utt_first_seg = r->first_seg_all.data() + b.first0;
  batch->utt_dst_row = r->dst_row_all.data() + b.u0;
  batch->first_ok_index = b.u0;
  ++r->next_batch;
  return XV_OK;
}

int xv_ark_reader_release(xv_ark_reader* r, int32_t slot) {
  if (!r) return fail(XV_EINVAL, "null argument");
  {
    std::lock_guard<std::mutex> lk(r->mu);
    if (r->released >= r->next_batch) return fail(XV_ESTATE, "no batch is out");
    if (slot != int32_t(r->released % int64_t(r->slot_buf.size()))) return fail(XV_ESTATE, "batches must be released in order");
    ++r->released;
  }
  r->cv_work.notify_all();
  return XV_OK;
}

void xv_ark_reader_close(xv_ark_reader* r) {
  if (!r) return;
  arkjob::stop_workers(r);
  for (size_t s = 0; s < r->slot_buf.size(); ++s) {
    float* p = r->slot_buf[s];
    if (!p) continue;
    if (!r->o.pinned) free(p);
    else if (!arkjob::pinned_pool().give(p, r->slot_bytes[s])) cudaFreeHost(p);
  }
  // tearing down the mapping of a large archive costs milliseconds (page tables, TLB shootdowns): not on the caller's clock
  if (r->map) {
    uint8_t* map = const_cast<uint8_t*>(r->map);
    const size_t size = size_t(r->file_size);
    const int fd = r->fd;
    try {
      std::thread([map, size, fd] { munmap(map, size); if (fd >= 0) close(fd); }).detach();
    } catch (...) {
      munmap(map, size);
      if (fd >= 0) close(fd);
    }
  } else if (r->fd >= 0) {
    close(r->fd);
  }
  delete r;
}

int xv_synth_mfcc(int device, float* out_dev, const int64_t* utt_id_host, const int32_t* len_host, int32_t n_utt, int32_t feat_dim,
                  uint64_t seed, void* stream_) {
  if (!out_dev || !utt_id_host || !len_host || n_utt < 1 || feat_dim < 1 || feat_dim > synth::MAX_DIM) return fail(XV_EINVAL, "bad argument");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  XV_CUDA(cudaSetDevice(device));
  // per-call scratch for the two small tables (freed stream-ordered)
  std::vector<int32_t> start(size_t(n_utt) + 1, 0);
  for (int i = 0; i < n_utt; ++i) {
    if (len_host[i] < 0) return fail(XV_EINVAL, "negative length");
    start[i + 1] = start[i] + len_host[i];
  }
  void* scratch = nullptr;
  const size_t id_bytes = size_t(n_utt) * 8, st_bytes = (size_t(n_utt) + 1) * 4;
  XV_CUDA(cudaMallocAsync(&scratch, id_bytes + st_bytes, stream));
  XV_CUDA(cudaMemcpyAsync(scratch, utt_id_host, id_bytes, cudaMemcpyHostToDevice, stream));
  XV_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(scratch) + id_bytes, start.data(), st_bytes, cudaMemcpyHostToDevice, stream));
  XV_CUDA(cudaStreamSynchronize(stream));            // `start` is a local: the copy must have left it
  synth::Args a{};
  a.out = out_dev;
  a.utt_id = static_cast<const int64_t*>(scratch);
  a.row_start = reinterpret_cast<const int32_t*>(static_cast<uint8_t*>(scratch) + id_bytes);
  a.n_utt = n_utt;
  a.feat_dim = feat_dim;
  a.total_rows = start[n_utt];
  a.seed = seed;
  // x has unit variance before the per-coefficient scale 12 / sqrt(1 + d) (synthetic.mfcc's law): Irwin-Hall(4) of 32-bit
  // uniforms has variance 2^64 / 3
  for (int d = 0; d < feat_dim; ++d)
    a.k[d] = float(std::sqrt(3.0) / 4294967296.0 * 12.0 / std::sqrt(1.0 + double(d)));
  if (a.total_rows > 0) {
    synth::synth_mfcc_kernel<<<unsigned((a.total_rows * feat_dim + 255) / 256), 256, 0, stream>>>(a);
    XV_CUDA(cudaGetLastError());
  }
  XV_CUDA(cudaFreeAsync(scratch, stream));
  return XV_OK;
}

int64_t xv_vec_ark_bytes(const int64_t* key_off, int64_t n, int32_t dim) {
  if (!key_off || n < 0 || dim < 0) return -1;
  return (key_off[n] - key_off[0]) + n * (11 + int64_t(dim) * 4);
}

int64_t xv_vec_ark_format(const char* key_blob, const int64_t* key_off, int64_t n, const float* vecs, int32_t dim, uint8_t* out,
                          int64_t out_cap, int64_t* marker_off, int32_t n_threads) {
  if (!key_off || n < 0 || dim < 0 || (n > 0 && (!key_blob || !vecs || !out))) return fail(XV_EINVAL, "bad argument");
  const int64_t total = xv_vec_ark_bytes(key_off, n, dim);
  if (total > out_cap) return fail(XV_ENOMEM, "output buffer too small");
  const int64_t per = 11 + int64_t(dim) * 4;
  auto run = [&](int64_t i0, int64_t i1) {
    for (int64_t i = i0; i < i1; ++i) {
      const int64_t klen = key_off[i + 1] - key_off[i];
      uint8_t* p = out + (key_off[i] - key_off[0]) + i * per;
      memcpy(p, key_blob + key_off[i], size_t(klen));
      p += klen;
      if (marker_off) marker_off[i] = (p + 1) - out;
      const uint8_t head[7] = {' ', 0, 'B', 'F', 'V', ' ', 4};
      memcpy(p, head, 7);
      const uint32_t d = uint32_t(dim);
      memcpy(p + 7, &d, 4);
      memcpy(p + 11, vecs + i * int64_t(dim), size_t(dim) * 4);
    }
  };
  const int nt = int(std::max<int64_t>(1, std::min<int64_t>(n_threads, n / 4096)));
  if (nt <= 1) run(0, n);
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back(run, n * t / nt, n * (t + 1) / nt);
    for (auto& x : th) x.join();
  }
  return total;
}

int64_t xv_scp_format(const char* key_blob, const int64_t* key_off, int64_t n, const char* ark_name, int64_t base,
                      const int64_t* marker_off, char* out, int64_t out_cap) {
  if (!key_off || !ark_name || !marker_off || n < 0 || (n > 0 && (!key_blob || !out))) return fail(XV_EINVAL, "bad argument");
  const size_t name_len = strlen(ark_name);
  int64_t pos = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t klen = key_off[i + 1] - key_off[i];
    if (pos + klen + int64_t(name_len) + 24 > out_cap) return fail(XV_ENOMEM, "output buffer too small");
    memcpy(out + pos, key_blob + key_off[i], size_t(klen));
    pos += klen;
    out[pos++] = ' ';
    memcpy(out + pos, ark_name, name_len);
    pos += int64_t(name_len);
    out[pos++] = ':';
    char digits[24];
    int nd = 0;
    uint64_t v = uint64_t(base + marker_off[i]);
    do { digits[nd++] = char('0' + v % 10); v /= 10; } while (v);
    while (nd) out[pos++] = digits[--nd];
    out[pos++] = '\n';
  }
  return pos;
}

}  // extern "C"
