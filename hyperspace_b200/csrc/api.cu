// api.cu -- the C ABI declared in include/hs_gpu.h.
#include <dirent.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <exception>
#include <map>
#include <mutex>
#include <thread>

#include "device_utils.cuh"
#include "engine.h"

using namespace hs;

#define HS_STR2(x) #x
#define HS_STR(x) HS_STR2(x)

namespace {

void set_err(char* err, size_t errlen, const char* msg) {
  if (err && errlen) {
    strncpy(err, msg, errlen - 1);
    err[errlen - 1] = 0;
  }
}

template <typename F>
int guarded(hs_ctx* ctx, char* err, size_t errlen, F&& f) {
  try {
    if (ctx) {
      cudaError_t e = cudaSetDevice(ctx->device);
      if (e != cudaSuccess) fail(HS_ECUDA, "cudaSetDevice(%d): %s", ctx->device, cudaGetErrorString(e));
      ctx->launches = 0;
    }
    f();
    return HS_OK;
  } catch (const hs::Error& e) {
    set_err(err, errlen, e.what());
    if (ctx) xfer_abort(ctx);
    cudaGetLastError();
    return e.code;
  } catch (const std::exception& e) {
    set_err(err, errlen, e.what());
    if (ctx) xfer_abort(ctx);
    cudaGetLastError();
    return HS_EINVAL;
  }
}

void write_file_atomic(const std::string& dir, const std::string& name, const uint8_t* data, uint64_t size) {
  const std::string tmp = dir + "/." + name + ".tmp";
  const std::string fin = dir + "/" + name;
  int fd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
  if (fd < 0) fail(HS_EIO, "cannot create %s", tmp.c_str());
  uint64_t done = 0;
  while (done < size) {
    ssize_t w = write(fd, data + done, size - done);
    if (w <= 0) {
      close(fd);
      unlink(tmp.c_str());
      fail(HS_EIO, "short write on %s", tmp.c_str());
    }
    done += (uint64_t)w;
  }
  close(fd);
  if (rename(tmp.c_str(), fin.c_str()) != 0) {
    unlink(tmp.c_str());
    fail(HS_EIO, "cannot rename %s", tmp.c_str());
  }
}

void mkdirs(const std::string& path) {
  std::string cur;
  for (size_t i = 0; i <= path.size(); i++) {
    if (i == path.size() || path[i] == '/') {
      if (!cur.empty()) mkdir(cur.c_str(), 0755);
    }
    if (i < path.size()) cur += path[i];
  }
}

// Spark's DataPathFilter (util/PathUtils.scala:34-39): names starting with '_' or '.' are not data files
void remove_data_files(const std::string& dir) {
  DIR* d = opendir(dir.c_str());
  if (!d) return;
  while (dirent* e = readdir(d)) {
    if (e->d_name[0] == '.' || e->d_name[0] == '_') continue;
    unlink((dir + "/" + e->d_name).c_str());
  }
  closedir(d);
}

void fill_lineage(hs_ctx* ctx, Table& t, const hs_source_file* files, int n_files);
void drop_deleted_rows(hs_ctx* ctx, Table& t, const int64_t* deleted, int ndeleted);

// gather every column of `t` through idx (n_out rows)
void gather_table(hs_ctx* ctx, Table& t, const uint32_t* d_idx, int64_t n_out) {
  for (DevColumn& c : t.cols) {
    Buf<uint8_t> nd(ctx, (size_t)n_out * c.width + 16);
    launch_gather_plain(ctx, c.data.get(), d_idx, n_out, c.width, nd.get());
    c.data = std::move(nd);
    if (c.valid) {
      Buf<uint8_t> nv(ctx, (size_t)n_out + 16);
      launch_gather_plain(ctx, c.valid.get(), d_idx, n_out, 1, nv.get());
      c.valid = std::move(nv);
    }
  }
  t.nrows = n_out;
}

__global__ void k_fill_u64(uint64_t* out, int64_t n, uint64_t v) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = v;
}
__global__ void k_fill_u32(uint32_t* out, int64_t n, uint32_t v) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = v;
}

// CoveringIndex.createIndexData lineage (index/covering/CoveringIndex.scala:152-186): every row carries the id of the
// source file it came from.
void fill_lineage(hs_ctx* ctx, Table& t, const hs_source_file* files, int n_files) {
  DevColumn dc;
  dc.name = "_data_file_id";
  dc.type = HS_TYPE_INT64;
  dc.width = 8;
  dc.schema.name = dc.name;
  dc.schema.type = pq::INT64;
  dc.schema.repetition = pq::OPTIONAL;
  dc.data.alloc(ctx, (size_t)t.nrows * 8 + 16);
  for (int f = 0; f < n_files; f++) {
    const int64_t b = t.file_row_begin[f], n = t.file_row_begin[f + 1] - b;
    if (n == 0) continue;
    const int grid = (int)std::min<int64_t>(ceil_div(n, 256), ctx->sm_count * 8);
    k_fill_u64<<<grid, 256, 0, ctx->stream>>>((uint64_t*)dc.data.get() + b, n, (uint64_t)files[f].file_id);
    HS_LAUNCH_CHECK(ctx);
  }
  t.cols.push_back(std::move(dc));
}

// Rows whose `_data_file_id` is in `deleted` are dropped (CoveringIndexTrait.refreshIncremental, deleted files branch,
// index/covering/CoveringIndexTrait.scala:78-94).
void drop_deleted_rows(hs_ctx* ctx, Table& t, const int64_t* deleted, int ndeleted) {
  int lc = -1;
  for (size_t c = 0; c < t.cols.size(); c++)
    if (t.cols[c].name == "_data_file_id") lc = (int)c;
  if (lc < 0) fail(HS_EINVAL, "deleted_file_ids given but the source has no _data_file_id column (index built without lineage)");
  const int64_t n = t.nrows;
  if (n == 0) return;
  Buf<uint32_t> mask(ctx, n);
  Buf<uint64_t> offs(ctx, n + 1);
  Buf<int64_t> d_del(ctx, ndeleted);
  copy_h2d(ctx, d_del.get(), deleted, sizeof(int64_t) * ndeleted);
  k_fill_u32<<<(int)std::min<int64_t>(ceil_div(n, 256), ctx->sm_count * 8), 256, 0, ctx->stream>>>(mask.get(), n, 1u);
  HS_LAUNCH_CHECK(ctx);
  launch_not_in_mask(ctx, (const int64_t*)t.cols[lc].data.get(), n, d_del.get(), ndeleted, mask.get());
  exclusive_scan_u32_u64(ctx, mask.get(), n, offs.get());
  uint64_t kept = 0;
  copy_d2h(ctx, &kept, offs.get() + n, 8);
  sync_stream(ctx);
  Buf<uint32_t> idx(ctx, std::max<uint64_t>(1, kept));
  launch_compact_indices(ctx, mask.get(), offs.get(), n, idx.get());
  gather_table(ctx, t, idx.get(), (int64_t)kept);
}

std::vector<std::string> names_of(const char* const* a, int na, const char* const* b, int nb) {
  std::vector<std::string> v;
  for (int i = 0; i < na; i++) v.emplace_back(a[i]);
  for (int i = 0; i < nb; i++) v.emplace_back(b[i]);
  return v;
}

// HS_OUT_FILES: the file images in res->h_arena become <out_dir>/<name>, all or nothing
// File-system traffic of the boundary (source files in, bucket files out) runs on a few host threads: one thread moves about
// 3-5 GB/s through the page cache, and the reference's writer runs one task per bucket in parallel as well
// (index/DataFrameWriterExtensions.scala:59-66 under Spark's FileFormatWriter).
int io_threads(size_t n_items) {
  const unsigned hw = std::thread::hardware_concurrency();
  size_t cap = 16;
  if (const char* e = getenv("HS_IO_THREADS")) cap = (size_t)std::max(1, atoi(e));
  return (int)std::max<size_t>(1, std::min<size_t>({cap, n_items, (size_t)std::max(1u, hw / 2)}));
}

// fn(i) for i in [0, n) on io_threads(n) threads; the first exception is rethrown on the caller's thread
template <typename Fn>
void parallel_files(size_t n, Fn fn) {
  const int nt = io_threads(n);
  if (nt <= 1) {
    for (size_t i = 0; i < n; i++) fn(i);
    return;
  }
  std::atomic<size_t> next{0};
  std::atomic<bool> failed{false};
  std::exception_ptr first;
  std::mutex mu;
  std::vector<std::thread> pool;
  for (int t = 0; t < nt; t++)
    pool.emplace_back([&] {
      for (;;) {
        const size_t i = next.fetch_add(1);
        if (i >= n || failed.load()) return;
        try {
          fn(i);
        } catch (...) {
          std::lock_guard<std::mutex> g(mu);
          if (!first) first = std::current_exception();
          failed.store(true);
          return;
        }
      }
    });
  for (auto& th : pool) th.join();
  if (first) std::rethrow_exception(first);
}

void write_result_files(hs_index_result* res, const std::string& dir, int save_mode) {
  if (dir.empty()) fail(HS_EINVAL, "HS_OUT_FILES needs out_dir");
  mkdirs(dir);
  if (save_mode == HS_SAVE_OVERWRITE) remove_data_files(dir);
  std::vector<char> written(res->files.size(), 0);
  try {
    parallel_files(res->files.size(), [&](size_t i) {
      const OutFile& f = res->files[i];
      write_file_atomic(dir, f.name, res->h_arena.get() + f.offset, f.size);
      written[i] = 1;
    });
  } catch (...) {
    for (size_t i = 0; i < written.size(); i++)
      if (written[i]) unlink((dir + "/" + res->files[i].name).c_str());  // all-or-nothing
    throw;
  }
  res->h_arena.release();
}

void finish_result(hs_ctx* ctx, EncodedFiles& enc, int output, const char* out_dir, int save_mode, hs_index_result* res,
                   hs_stats* st) {
  res->ctx = ctx;
  res->output = output;
  res->files = enc.files;
  if (output == HS_OUT_DEVICE) {
    res->d_arena = std::move(enc.arena);
    return;
  }
  StageTimer t(ctx);
  t.start();
  res->h_arena.alloc(ctx, std::max<uint64_t>(enc.arena_bytes, 16), /*pinned=*/true);
  if (enc.arena_bytes)
    copy_d2h(ctx, res->h_arena.get(), enc.arena.get(), enc.arena_bytes);
  t.stop();
  sync_stream(ctx);
  st->ms_d2h += t.ms();
  if (output == HS_OUT_FILES) write_result_files(res, out_dir ? out_dir : "", save_mode);
}

void read_file_into(const char* path, uint8_t* dst, uint64_t size) {
  int fd = open(path, O_RDONLY);
  if (fd < 0) fail(HS_EIO, "cannot open %s", path);
  uint64_t got = 0;
  while (got < size) {
    ssize_t r = read(fd, dst + got, size - got);
    if (r <= 0) {
      close(fd);
      fail(HS_EIO, "short read on %s", path);
    }
    got += (uint64_t)r;
  }
  close(fd);
}

}  // namespace

// =====================================================================================================================
extern "C" {

int hs_abi_version(void) { return HS_ABI_VERSION; }

const char* hs_build_info(void) {
  return "hyperspace_b200 libhs_gpu 0.1.0; arch sm_100a; CUDA " HS_STR(CUDART_VERSION) "; nvcc -O3 -lineinfo";
}

int hs_init(int device_id, void* cuda_stream, hs_ctx** out, char* err, size_t errlen) {
  if (!out) return HS_EINVAL;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    char buf[256];
    snprintf(buf, sizeof buf, "no CUDA device available (%s); libhs_gpu has no CPU fallback",
             e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    set_err(err, errlen, buf);
    cudaGetLastError();
    return HS_ENODEVICE;
  }
  if (device_id < 0 || device_id >= count) {
    set_err(err, errlen, "device id out of range");
    return HS_EINVAL;
  }
  hs_ctx* ctx = new hs_ctx();
  ctx->device = device_id;
  int rc = guarded(nullptr, err, errlen, [&] {
    HS_CUDA(cudaSetDevice(device_id));
    cudaDeviceProp prop;
    HS_CUDA(cudaGetDeviceProperties(&prop, device_id));
    ctx->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) fail(HS_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device_id, prop.major, prop.minor);
    if (cuda_stream) {
      ctx->stream = (cudaStream_t)cuda_stream;
      ctx->own_stream = false;
    } else {
      HS_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
      ctx->own_stream = true;
    }
    HS_CUDA(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
    HS_CUDA(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
  });
  if (rc != HS_OK) {
    delete ctx;
    return rc == HS_ECUDA ? HS_ENODEVICE : rc;
  }
  *out = ctx;
  return HS_OK;
}

void hs_shutdown(hs_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  xfer_release(ctx);
  if (ctx->decode_cache && ctx->decode_cache_free) ctx->decode_cache_free(ctx->decode_cache);
  ctx->decode_cache = nullptr;
  comm_destroy(ctx);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  if (ctx->h2d_stream) {
    cudaStreamSynchronize(ctx->h2d_stream);
    cudaStreamDestroy(ctx->h2d_stream);
  }
  if (ctx->d2h_stream) {
    cudaStreamSynchronize(ctx->d2h_stream);
    cudaStreamDestroy(ctx->d2h_stream);
  }
  delete ctx;
}

void hs_trim(hs_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->pool.trim();
}

void hs_profile_enable(hs_ctx* ctx, int on) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& k : ctx->kevents) {
    ctx->event_pool.push_back(k.a);
    ctx->event_pool.push_back(k.b);
  }
  ctx->kevents.clear();
  ctx->profile = on != 0;
}

int hs_profile_report(hs_ctx* ctx, char* out, size_t outlen) {
  if (!ctx || !out || !outlen) return HS_EINVAL;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  std::map<std::string, std::pair<int, double>> acc;
  for (auto& k : ctx->kevents) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, k.a, k.b) != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    auto& e = acc[k.name];
    e.first++;
    e.second += ms;
  }
  std::string s = "{";
  bool first = true;
  for (auto& kv : acc) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s\"%s\": {\"launches\": %d, \"ms\": %.6f}", first ? "" : ", ", kv.first.c_str(), kv.second.first,
             kv.second.second);
    s += buf;
    first = false;
  }
  s += "}";
  if (s.size() + 1 > outlen) return HS_ENOMEM;
  memcpy(out, s.c_str(), s.size() + 1);
  for (auto& k : ctx->kevents) {
    ctx->event_pool.push_back(k.a);
    ctx->event_pool.push_back(k.b);
  }
  ctx->kevents.clear();
  return HS_OK;
}

void* hs_host_alloc(hs_ctx* ctx, size_t bytes) {
  if (!ctx) return nullptr;
  try {
    cudaSetDevice(ctx->device);
    return ctx->pool.get(bytes, true);
  } catch (...) {
    return nullptr;
  }
}

void hs_host_free(hs_ctx* ctx, void* p) {
  if (ctx && p) ctx->pool.put(p);
}

// ---- staging: host file images -> device, asynchronously on the H2D copy stream -----------------------------------------

int hs_stage_sources(hs_ctx* ctx, const hs_source_file* files, int32_t n_files, hs_staged** out, char* err, size_t errlen) {
  if (!ctx || !out || n_files < 0 || (n_files > 0 && !files)) return HS_EINVAL;
  *out = nullptr;
  std::unique_ptr<hs_staged> sg(new hs_staged());
  sg->ctx = ctx;
  int rc = guarded(ctx, err, errlen, [&] {
    const int launches_before = ctx->launches;
    std::vector<uint64_t> sizes(n_files), off(n_files);
    uint64_t total = 0;
    for (int f = 0; f < n_files; f++) {
      const hs_source_file& sf = files[f];
      if (sf.data && sf.on_device) fail(HS_EINVAL, "source file %d is already on the device", f);
      if (sf.data) {
        sizes[f] = sf.size;
      } else {
        if (!sf.path) fail(HS_EINVAL, "source file %d has neither data nor path", f);
        struct stat st;
        if (stat(sf.path, &st) != 0) fail(HS_EIO, "cannot stat %s", sf.path);
        sizes[f] = (uint64_t)st.st_size;
      }
      if (sizes[f] < 12) fail(HS_EFORMAT, "source file %d: too small to be a Parquet file", f);
      off[f] = total;
      total += round_up(sizes[f], 16) + 16;
    }
    sg->d_images.alloc(ctx, std::max<uint64_t>(total, 16));
    HS_CUDA(cudaEventCreate(&sg->begin));
    HS_CUDA(cudaEventRecord(sg->begin, ctx->h2d_stream));
    sg->files.resize(n_files);
    sg->names.resize(n_files);
    sg->metas.resize(n_files);
    // file system sources are read into pinned memory first (kept until the copy has completed): the buffers come from the
    // pool on this thread, the reads run on a few threads, and each file's H2D copy is queued as soon as its read is done
    std::vector<uint8_t*> pinned(n_files, nullptr);
    std::vector<size_t> disk;
    for (int f = 0; f < n_files; f++)
      if (!files[f].data) {
        sg->staging.emplace_back(ctx, sizes[f], /*pinned=*/true);
        pinned[f] = sg->staging.back().get();
        disk.push_back((size_t)f);
      }
    std::vector<std::atomic<int>> read_done(n_files);
    for (auto& d : read_done) d.store(0);
    std::exception_ptr read_error;
    std::thread reader;
    if (!disk.empty())
      reader = std::thread([&] {
        try {
          parallel_files(disk.size(), [&](size_t j) {
            const size_t f = disk[j];
            read_file_into(files[f].path, pinned[f], sizes[f]);
            read_done[f].store(1, std::memory_order_release);
          });
        } catch (...) {
          read_error = std::current_exception();
        }
        for (size_t f : disk)  // wake the consumer whatever happened
          if (!read_done[f].load()) read_done[f].store(-1, std::memory_order_release);
      });
    struct Joiner {
      std::thread& t;
      ~Joiner() {
        if (t.joinable()) t.join();
      }
    } joiner{reader};
    for (int f = 0; f < n_files; f++) {
      const hs_source_file& sf = files[f];
      sg->names[f] = sf.path ? sf.path : ("<memory file " + std::to_string(f) + ">");
      const uint8_t* host = (const uint8_t*)sf.data;
      if (!host) {
        int st;
        while ((st = read_done[f].load(std::memory_order_acquire)) == 0) std::this_thread::yield();
        if (st < 0) {
          if (reader.joinable()) reader.join();
          if (read_error) std::rethrow_exception(read_error);
          fail(HS_EIO, "cannot read %s", sf.path);
        }
        host = pinned[f];
      }
      // the footer is parsed here, from host memory, so that the build never has to fetch it back from the device
      sg->metas[f] = std::make_shared<hs::pq::FileMeta>(hs::pq::parse_footer(host, sizes[f], sg->names[f].c_str()));
      HS_CUDA(cudaMemcpyAsync(sg->d_images.get() + off[f], host, sizes[f], cudaMemcpyHostToDevice, ctx->h2d_stream));
      hs_source_file& o = sg->files[f];
      o.path = nullptr;  // patched to names[f].c_str() by hs_staged_file (the vector may still move here)
      o.data = sg->d_images.get() + off[f];
      o.size = sizes[f];
      o.file_id = sf.file_id;
      o.on_device = 1;
      o.reserved = 0;
      sg->bytes += sizes[f];
    }
    HS_CUDA(cudaEventCreate(&sg->ready));
    HS_CUDA(cudaEventRecord(sg->ready, ctx->h2d_stream));
    for (int f = 0; f < n_files; f++) ctx->staged[sg->files[f].data] = hs_ctx::StagedImage{sg->metas[f], sg->ready};
    ctx->launches = launches_before;
  });
  if (rc == HS_OK) *out = sg.release();
  return rc;
}

int32_t hs_staged_num_files(const hs_staged* s) { return s ? (int32_t)s->files.size() : 0; }

int hs_staged_file(const hs_staged* s, int32_t i, hs_source_file* out) {
  if (!s || !out || i < 0 || i >= (int32_t)s->files.size()) return HS_EINVAL;
  *out = s->files[i];
  out->path = s->names[i].c_str();
  return HS_OK;
}

// HS_TIMELINE=1: begin / end of every staged copy, build and drain relative to the first event seen, on stderr
static void timeline(hs_ctx* ctx, const char* what, cudaEvent_t a, cudaEvent_t b) {
  static const bool on = getenv("HS_TIMELINE") != nullptr;
  if (!on || !a || !b) return;
  static cudaEvent_t base = nullptr;
  if (!base) base = a;  // (leaks one reference to an event that may be destroyed later: diagnostics only, first event kept alive)
  float t0 = 0, t1 = 0;
  if (cudaEventElapsedTime(&t0, base, a) != cudaSuccess || cudaEventElapsedTime(&t1, base, b) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  fprintf(stderr, "[hs timeline] %-8s %10.2f -> %10.2f  (%8.2f ms)\n", what, t0, t1, t1 - t0);
}

int hs_staged_wait(hs_staged* s, float* ms_copy) {
  if (!s) return HS_EINVAL;
  cudaSetDevice(s->ctx->device);
  if (cudaEventSynchronize(s->ready) != cudaSuccess) return HS_ECUDA;
  if (ms_copy && cudaEventElapsedTime(ms_copy, s->begin, s->ready) != cudaSuccess) *ms_copy = 0;
  timeline(s->ctx, "H2D", s->begin, s->ready);
  return HS_OK;
}

void hs_staged_free(hs_staged* s) {
  if (!s) return;
  hs_ctx* ctx = s->ctx;
  cudaSetDevice(ctx->device);
  if (s->ready) {
    cudaEventSynchronize(s->ready);  // the copies read caller memory / pinned staging and write d_images
    cudaEventDestroy(s->ready);
  }
  if (s->begin && !getenv("HS_TIMELINE")) cudaEventDestroy(s->begin);  // (the timeline's base event is one of these)
  for (const hs_source_file& f : s->files) ctx->staged.erase(f.data);
  cudaStreamSynchronize(ctx->stream);  // a build that decodes these images may still be running
  delete s;
}

// ---- createIndex -----------------------------------------------------------------------------------------------------

int hs_create_index_async(hs_ctx* ctx, const hs_index_spec* spec, hs_pending** out, char* err, size_t errlen) {
  if (!ctx || !spec || !out) return HS_EINVAL;
  *out = nullptr;
  std::unique_ptr<hs_pending> pd(new hs_pending());
  pd->ctx = ctx;
  pd->res.reset(new hs_index_result());
  hs_stats& st = pd->st;
  memset(&st, 0, sizeof st);
  hs_index_result* res = pd->res.get();
  int rc = guarded(ctx, err, errlen, [&] {
    if (spec->n_indexed < 1) fail(HS_EINVAL, "at least one indexed column is required");
    if (spec->n_files < 0 || (spec->n_files > 0 && !spec->files)) fail(HS_EINVAL, "bad source file list");
    if (spec->output == HS_OUT_FILES && !spec->out_dir) fail(HS_EINVAL, "HS_OUT_FILES needs out_dir");
    HS_CUDA(cudaEventCreate(&pd->t_begin));
    HS_CUDA(cudaEventCreate(&pd->t_compute_end));
    HS_CUDA(cudaEventRecord(pd->t_begin, ctx->stream));
    std::vector<std::string> cols = names_of(spec->indexed_columns, spec->n_indexed, spec->included_columns, spec->n_included);
    for (size_t i = 0; i < cols.size(); i++)
      for (size_t j = i + 1; j < cols.size(); j++)
        if (cols[i] == cols[j]) fail(HS_EINVAL, "duplicate column '%s' in index config", cols[i].c_str());
    EncodedFiles enc;
    {
      Table table;
      // Included columns whose source pages are all dictionary-encoded travel through the build as 16-bit dictionary codes
      // (late materialisation).  Only where nothing between decode and encode needs their values: the fused partition (on
      // one GPU, or writing straight into the owners' memory on several), no rows to drop.  HS_NO_CARRY=1 switches it off
      // (A/B measurements; on several GPUs every rank must be given the same setting).
      const bool no_carry = getenv("HS_NO_CARRY") != nullptr;
      CarryOptions carry;
      if ((ctx->world == 1 || p2p_exchange_supported(ctx, spec->num_buckets)) && !spec->disable_dictionary &&
          spec->n_deleted_file_ids == 0 && !no_carry && fused_partition_supported(spec->num_buckets)) {
        carry.first_col = spec->n_indexed;
        carry.num_segments = spec->num_buckets;
      }
      // PLAIN, null-free, value-aligned columns are not decoded either: hash and partition read them in place from the file
      // images (zero copy), which therefore stay alive until the rows have been partitioned.  HS_NO_ZEROCOPY=1: A/B switch.
      if (fused_partition_supported(spec->num_buckets) && spec->n_deleted_file_ids == 0 && !getenv("HS_NO_ZEROCOPY")) {
        carry.zc_tile_rows = fused_tile_rows(p2p_exchange_supported(ctx, spec->num_buckets));
        carry.zc_first_col = spec->n_indexed;
        carry.zc_key = spec->n_indexed == 1;
      }
      SourceSet src;
      open_sources(ctx, spec->files, spec->n_files, &src, &st);
      decode_sources(ctx, src, cols, nullptr, &table, &st, &carry);
      const bool has_strings = table.has_strings;
      if (has_strings && ctx->world > 1)
        fail(HS_EUNSUPPORTED, "string / binary columns are not exchanged between GPUs yet: build this index on one GPU");
      if (spec->lineage) fill_lineage(ctx, table, spec->files, spec->n_files);
      if (spec->n_deleted_file_ids > 0) drop_deleted_rows(ctx, table, spec->deleted_file_ids, spec->n_deleted_file_ids);
      IndexedRows rows;
      if (p2p_exchange_supported(ctx, spec->num_buckets)) {
        // partition + exchange fused over NVLink peer memory, then the local sort
        exchange_partition_p2p(ctx, table, spec->n_indexed, spec->num_buckets, &rows, &st);
        sort_partitioned_rows(ctx, spec->n_indexed, spec->num_buckets, &rows, &st, /*defer_settle=*/true);
      } else {
        if (ctx->world > 1) exchange_rows(ctx, table, spec->n_indexed, spec->num_buckets, &st);  // NCCL all-to-all
        index_rows(ctx, table, spec->n_indexed, spec->num_buckets, &rows, &st, /*defer_settle=*/true);
      }
      // From here to the end of encode_segments the host does not wait for the GPU unless it has to: the sort kernels are
      // queued, the verdict of the tie fix-up is on its way (settle_sort below), and the encoder lays the pages out on the
      // host while the rows are still being sorted.
      if (!has_strings) src.release_images();  // every column is materialised bucket-major now (string
                                                         // references keep pointing into the images until the encode)

      EncodeRequest req;
      req.table = &rows.part;
      req.d_perm = rows.sorted_perm;
      req.d_sorted_keys = rows.sorted_keys;
      req.plan = &rows.plan;
      req.seg_offsets = rows.bucket_offsets;
      req.rows_per_page = spec->rows_per_page;
      req.rows_per_row_group = spec->rows_per_row_group;
      req.use_dictionary = spec->disable_dictionary == 0;
      if (spec->compression != HS_CODEC_UNCOMPRESSED && spec->compression != HS_CODEC_SNAPPY)
        fail(HS_EUNSUPPORTED, "compression codec %d; the GPU path writes UNCOMPRESSED and SNAPPY pages", spec->compression);
      req.codec = spec->compression == HS_CODEC_SNAPPY ? pq::SNAPPY : pq::UNCOMPRESSED;
      const std::string uuid = spec->job_uuid ? spec->job_uuid : make_uuid();
      req.seg_names.resize(spec->num_buckets);
      for (int b = 0; b < spec->num_buckets; b++) {
        // Spark FileFormatWriter: part-<task>-<jobUUID>_<bucket>.c000<codec ext>.parquet ; bucket id parsed back by
        // BucketingUtils.getBucketId (relied on by actions/OptimizeAction.scala:110)
        char nm[160];
        snprintf(nm, sizeof nm, "part-%05d-%s_%05d.c000%s.parquet", b, uuid.c_str(), b, req.codec == pq::SNAPPY ? ".snappy" : "");
        req.seg_names[b] = nm;
      }
      req.probe = rows.probe.get();
      encode_segments(ctx, req, &enc, &st);  // synchronises the stream before it returns
      if (settle_sort(ctx, &rows, &st)) {    // (rare) the fix-up gave up on a long run of equal key prefixes: the rows were sorted
        req.probe = nullptr;                 // again with full passes, so the pages are gathered again
        req.d_perm = rows.sorted_perm;
        req.d_sorted_keys = rows.sorted_keys;
        enc = EncodedFiles();
        encode_segments(ctx, req, &enc, &st);
      }
      st.rows_out = rows.part.nrows;
      // the decoded / partitioned / sorted intermediates go back to the pool here: while this call's index files drain to
      // the host, the next call can already build in the same memory
    }
    HS_CUDA(cudaEventRecord(pd->t_compute_end, ctx->stream));
    res->ctx = ctx;
    res->output = spec->output;
    res->files = enc.files;
    pd->save_mode = spec->save_mode;
    if (spec->out_dir) pd->out_dir = spec->out_dir;
    st.gpu_launches = ctx->launches;
    if (spec->output == HS_OUT_DEVICE) {
      res->d_arena = std::move(enc.arena);
      return;
    }
    // device -> host on the D2H copy stream; hs_pending_wait picks it up
    pd->d_arena = std::move(enc.arena);
    res->h_arena.alloc(ctx, std::max<uint64_t>(enc.arena_bytes, 16), /*pinned=*/true);
    HS_CUDA(cudaEventCreate(&pd->t_d2h_begin));
    HS_CUDA(cudaEventCreate(&pd->t_d2h_end));
    HS_CUDA(cudaStreamWaitEvent(ctx->d2h_stream, pd->t_compute_end, 0));
    HS_CUDA(cudaEventRecord(pd->t_d2h_begin, ctx->d2h_stream));
    if (enc.arena_bytes)
      HS_CUDA(cudaMemcpyAsync(res->h_arena.get(), pd->d_arena.get(), enc.arena_bytes, cudaMemcpyDeviceToHost, ctx->d2h_stream));
    HS_CUDA(cudaEventRecord(pd->t_d2h_end, ctx->d2h_stream));
    pd->has_d2h = true;
  });
  if (rc == HS_OK) *out = pd.release();
  return rc;
}

int hs_pending_wait(hs_pending* p, hs_index_result** out, hs_stats* stats, char* err, size_t errlen) {
  if (!p || !out) return HS_EINVAL;
  *out = nullptr;
  std::unique_ptr<hs_pending> pd(p);  // consumed either way
  hs_ctx* ctx = pd->ctx;
  const int launches = ctx->launches;
  int rc = guarded(ctx, err, errlen, [&] {
    hs_stats& st = pd->st;
    float ms = 0;
    if (pd->has_d2h) {
      HS_CUDA(cudaEventSynchronize(pd->t_d2h_end));
      HS_CUDA(cudaEventElapsedTime(&ms, pd->t_d2h_begin, pd->t_d2h_end));
      st.ms_d2h += ms;
      HS_CUDA(cudaEventElapsedTime(&ms, pd->t_begin, pd->t_d2h_end));
      st.ms_total = ms;
      timeline(ctx, "build", pd->t_begin, pd->t_compute_end);
      timeline(ctx, "D2H", pd->t_d2h_begin, pd->t_d2h_end);
      pd->d_arena.release();
    } else {
      HS_CUDA(cudaEventSynchronize(pd->t_compute_end));
      HS_CUDA(cudaEventElapsedTime(&ms, pd->t_begin, pd->t_compute_end));
      st.ms_total = ms;
    }
    if (pd->res->output == HS_OUT_FILES) {
      const auto w0 = std::chrono::steady_clock::now();
      write_result_files(pd->res.get(), pd->out_dir, pd->save_mode);
      st.ms_write += std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - w0).count();
    }
  });
  ctx->launches = launches;
  if (stats) *stats = pd->st;
  if (rc == HS_OK) *out = pd->res.release();
  return rc;
}

void hs_pending_cancel(hs_pending* p) {
  if (!p) return;
  cudaSetDevice(p->ctx->device);
  if (p->has_d2h) cudaEventSynchronize(p->t_d2h_end);
  delete p;
}

int hs_create_index(hs_ctx* ctx, const hs_index_spec* spec, hs_index_result** out, hs_stats* stats, char* err,
                    size_t errlen) {
  if (!ctx || !spec || !out) return HS_EINVAL;
  *out = nullptr;
  if (stats) memset(stats, 0, sizeof *stats);
  hs_pending* pd = nullptr;
  int rc = hs_create_index_async(ctx, spec, &pd, err, errlen);
  if (rc != HS_OK) return rc;
  const int launches = ctx->launches;
  rc = hs_pending_wait(pd, out, stats, err, errlen);
  if (stats) stats->gpu_launches = launches;
  return rc;
}

int32_t hs_result_num_files(const hs_index_result* r) { return r ? (int32_t)r->files.size() : 0; }

int hs_result_file(const hs_index_result* r, int32_t i, int32_t* bucket, const char** name, const void** data,
                   uint64_t* size, int64_t* rows) {
  if (!r || i < 0 || i >= (int32_t)r->files.size()) return HS_EINVAL;
  const OutFile& f = r->files[i];
  if (bucket) *bucket = f.bucket;
  if (name) *name = f.name.c_str();
  if (data) {
    if (r->output == HS_OUT_DEVICE) *data = r->d_arena.get() + f.offset;
    else if (r->output == HS_OUT_HOST) *data = r->h_arena.get() + f.offset;
    else *data = nullptr;
  }
  if (size) *size = f.size;
  if (rows) *rows = f.rows;
  return HS_OK;
}

void hs_result_free(hs_index_result* r) {
  if (!r) return;
  if (r->ctx) {
    cudaSetDevice(r->ctx->device);
    cudaStreamSynchronize(r->ctx->stream);
  }
  delete r;
}

// ---------------------------------------------------------------------------------------------------------------------
// read side
// ---------------------------------------------------------------------------------------------------------------------

static void batch_from_gather(hs_ctx* ctx, const Table& t, const std::vector<int>& col_idx, const uint32_t* d_idx,
                              int64_t n_out, hs_batch* b) {
  // all gathers first, then all copies, one synchronisation at the end
  std::vector<Buf<uint8_t>> d_data, d_valid;
  std::vector<Buf<uint64_t>> d_offsets(col_idx.size());
  std::vector<uint64_t> str_bytes(col_idx.size(), 0);
  for (size_t k = 0; k < col_idx.size(); k++) {
    const int ci = col_idx[k];
    const DevColumn& c = t.cols[ci];
    if (c.type == HS_TYPE_STRING) {
      // values are references into the source images (still alive here): lengths -> offsets -> one copy of the bytes
      const uint8_t* valid = c.has_nulls ? c.valid.get() : nullptr;
      Buf<uint32_t> lens(ctx, (size_t)std::max<int64_t>(1, n_out));
      d_offsets[k].alloc(ctx, (size_t)n_out + 1);
      launch_string_lengths(ctx, (const uint64_t*)c.data.get(), valid, d_idx, n_out, lens.get());
      exclusive_scan_u32_u64(ctx, lens.get(), n_out, d_offsets[k].get());
      copy_d2h(ctx, &str_bytes[k], d_offsets[k].get() + n_out, 8);
      sync_stream(ctx);
      d_data.emplace_back(ctx, std::max<uint64_t>(1, str_bytes[k]));
      launch_copy_strings(ctx, (const uint64_t*)c.data.get(), valid, d_idx, n_out, d_offsets[k].get(), d_data.back().get());
      d_valid.emplace_back();
      if (c.has_nulls) {
        d_valid.back().alloc(ctx, (size_t)std::max<int64_t>(1, n_out));
        launch_gather_plain(ctx, c.valid.get(), d_idx, n_out, 1, d_valid.back().get());
      }
      continue;
    }
    d_data.emplace_back(ctx, (size_t)std::max<int64_t>(1, n_out) * c.width);
    launch_gather_plain(ctx, c.data.get(), d_idx, n_out, c.width, d_data.back().get());
    d_valid.emplace_back();
    if (c.has_nulls) {
      d_valid.back().alloc(ctx, (size_t)std::max<int64_t>(1, n_out));
      launch_gather_plain(ctx, c.valid.get(), d_idx, n_out, 1, d_valid.back().get());
    }
  }
  for (size_t i = 0; i < col_idx.size(); i++) {
    const DevColumn& c = t.cols[col_idx[i]];
    hs_batch::Col bc;
    bc.name = c.name;
    bc.type = c.type;
    bc.has_valid = c.has_nulls;
    const bool is_str = c.type == HS_TYPE_STRING;
    const size_t data_bytes = is_str ? (size_t)str_bytes[i] : (size_t)n_out * c.width;
    bc.total_bytes = is_str ? str_bytes[i] : 0;
    if (b->on_device) {  // the next GPU operator consumes the columns where they are
      bc.data = std::move(d_data[i]);
      if (is_str) bc.offsets = std::move(d_offsets[i]);
      if (c.has_nulls) bc.valid = std::move(d_valid[i]);
    } else {
      bc.data.alloc(ctx, std::max<size_t>(1, data_bytes), true);
      if (data_bytes) copy_d2h(ctx, bc.data.get(), d_data[i].get(), data_bytes);
      if (is_str) {
        bc.offsets.alloc(ctx, (size_t)n_out + 1, true);
        copy_d2h(ctx, bc.offsets.get(), d_offsets[i].get(), 8 * ((size_t)n_out + 1));
      }
      if (c.has_nulls) {
        bc.valid.alloc(ctx, (size_t)std::max<int64_t>(1, n_out), true);
        if (n_out) copy_d2h(ctx, bc.valid.get(), d_valid[i].get(), (size_t)n_out);
      }
    }
    b->cols.push_back(std::move(bc));
  }
  sync_stream(ctx);
  b->nrows = n_out;
}

// int32 keys are widened once so that the comparison kernels (range bounds, predicate mask, join probes) stay int64-only
__global__ void k_widen_i32(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i];
}

static const int64_t* widened_key(hs_ctx* ctx, const DevColumn& c, int64_t n, Buf<int64_t>* hold) {
  if (c.type == HS_TYPE_INT64) return (const int64_t*)c.data.get();
  if (c.type != HS_TYPE_INT32) fail(HS_EUNSUPPORTED, "key column '%s' must be int32 or int64 on the read side", c.name.c_str());
  hold->alloc(ctx, std::max<int64_t>(1, n));
  if (n) {
    k_widen_i32<<<(int)std::min<int64_t>(ceil_div(n, 256), ctx->sm_count * 16), 256, 0, ctx->stream>>>((const int32_t*)c.data.get(), n,
                                                                                                      hold->get());
    HS_LAUNCH_CHECK(ctx);
  }
  return hold->get();
}

__global__ void k_ranges_to_indices(const int64_t* __restrict__ bounds, const uint64_t* __restrict__ seg_offsets,
                                    const uint64_t* __restrict__ out_offsets, int nseg, uint32_t* __restrict__ out_idx) {
  // one CTA per segment
  const int s = blockIdx.x;
  if (s >= nseg) return;
  const int64_t first = bounds[2 * s], last = bounds[2 * s + 1];
  const uint64_t base = seg_offsets[s], o = out_offsets[s];
  for (int64_t i = first + threadIdx.x; i < last; i += blockDim.x) out_idx[o + (i - first)] = (uint32_t)(base + i);
}

int hs_filter_scan(hs_ctx* ctx, const hs_scan_spec* spec, hs_batch** out, hs_stats* stats, char* err, size_t errlen) {
  if (!ctx || !spec || !out) return HS_EINVAL;
  *out = nullptr;
  hs_stats st;
  memset(&st, 0, sizeof st);
  std::unique_ptr<hs_batch> res(new hs_batch());
  res->ctx = ctx;
  res->on_device = spec->output == HS_OUT_DEVICE;
  int rc = guarded(ctx, err, errlen, [&] {
    StageTimer total(ctx);
    total.start();
    if (!spec->key_column) fail(HS_EINVAL, "key_column is required");
    // columns to decode: key first, then the projection, then lineage when deletes must be filtered
    std::vector<std::string> cols{spec->key_column};
    std::vector<int> proj_idx;
    for (int i = 0; i < spec->n_projected; i++) {
      std::string nm = spec->projected_columns[i];
      auto it = std::find(cols.begin(), cols.end(), nm);
      if (it == cols.end()) {
        cols.push_back(nm);
        proj_idx.push_back((int)cols.size() - 1);
      } else {
        proj_idx.push_back((int)(it - cols.begin()));
      }
    }
    int lineage_col = -1;
    if (spec->n_deleted_file_ids > 0) {
      auto it = std::find(cols.begin(), cols.end(), std::string("_data_file_id"));
      if (it == cols.end()) {
        cols.push_back("_data_file_id");
        lineage_col = (int)cols.size() - 1;
      } else {
        lineage_col = (int)(it - cols.begin());
      }
    }
    const bool try_sorted = spec->sorted_on_key && spec->n_deleted_file_ids == 0;
    SourceSet src;
    open_sources(ctx, spec->files, spec->n_files, &src, &st);
    Table t;
    if (try_sorted) {
      // phase 1: only the key column; the other columns are decoded after the binary search, restricted to the pages
      // that hold qualifying rows
      decode_sources(ctx, src, {cols[0]}, nullptr, &t, &st);
    } else {
      decode_sources(ctx, src, cols, nullptr, &t, &st);
    }
    const bool str_key = t.cols[0].type == HS_TYPE_STRING;
    if (!str_key && t.cols[0].type != HS_TYPE_INT64 && t.cols[0].type != HS_TYPE_INT32)
      fail(HS_EUNSUPPORTED, "filter scan: key column must be int32 / int64 / string");
    const int64_t n = t.nrows;
    Buf<int64_t> k64;
    const int64_t* d_keys = str_key ? nullptr : widened_key(ctx, t.cols[0], n, &k64);
    // string bounds: device copies of the bytes, addressed like every string value (a reference)
    Buf<uint8_t> d_bound_bytes;
    uint64_t lo_ref = 0, hi_ref = 0;
    if (str_key) {
      if ((spec->has_lo && spec->lo_len && !spec->lo_bytes) || (spec->has_hi && spec->hi_len && !spec->hi_bytes))
        fail(HS_EINVAL, "filter scan: string key '%s' needs lo_bytes / hi_bytes", t.cols[0].name.c_str());
      if (spec->lo_len > kMaxStringLen || spec->hi_len > kMaxStringLen) fail(HS_EUNSUPPORTED, "string bound longer than 65535 bytes");
      const uint32_t ll = spec->has_lo ? spec->lo_len : 0, hl = spec->has_hi ? spec->hi_len : 0;
      d_bound_bytes.alloc(ctx, (size_t)ll + hl + 16);
      if (ll) copy_h2d(ctx, d_bound_bytes.get(), spec->lo_bytes, ll);
      if (hl) copy_h2d(ctx, d_bound_bytes.get() + ll, spec->hi_bytes, hl);
      lo_ref = string_ref(d_bound_bytes.get(), ll);
      hi_ref = string_ref(d_bound_bytes.get() + ll, hl);
    }
    StageTimer t_scan(ctx);
    t_scan.start();
    Buf<uint32_t> idx;
    int64_t n_out = 0;
    bool sorted = try_sorted && !t.cols[0].has_nulls;
    if (try_sorted && !sorted) {  // nulls in the key: fall back to the predicate scan over all columns
      Table full;
      decode_sources(ctx, src, cols, nullptr, &full, &st);
      t = std::move(full);
      if (!str_key) d_keys = widened_key(ctx, t.cols[0], n, &k64);
    }
    if (sorted) {
      // K7: two binary searches per file
      const int nseg = spec->n_files;
      std::vector<uint64_t> seg(nseg + 1);
      for (int f = 0; f <= nseg; f++) seg[f] = (uint64_t)t.file_row_begin[f];
      Buf<uint64_t> d_seg(ctx, nseg + 1);
      Buf<int64_t> d_bounds(ctx, 2 * std::max(1, nseg));
      copy_h2d(ctx, d_seg.get(), seg.data(), 8 * (nseg + 1));
      if (str_key)
        launch_range_bounds_strings(ctx, (const uint64_t*)t.cols[0].data.get(), d_seg.get(), nseg, spec->has_lo, lo_ref, spec->has_hi,
                                    hi_ref, d_bounds.get());
      else
        launch_range_bounds(ctx, d_keys, d_seg.get(), nseg, spec->has_lo, spec->lo, spec->has_hi, spec->hi, d_bounds.get());
      std::vector<int64_t> bounds(2 * std::max(1, nseg));
      copy_d2h(ctx, bounds.data(), d_bounds.get(), 16 * nseg);
      sync_stream(ctx);
      std::vector<uint64_t> oo(nseg + 1, 0);
      for (int f = 0; f < nseg; f++) oo[f + 1] = oo[f] + (uint64_t)(bounds[2 * f + 1] - bounds[2 * f]);
      if (cols.size() > 1) {  // phase 2: decode the other columns, only the pages inside each file's [first, last)
        std::vector<std::pair<int64_t, int64_t>> windows(nseg);
        for (int f = 0; f < nseg; f++) windows[f] = {bounds[2 * f], bounds[2 * f + 1]};
        std::vector<std::string> rest(cols.begin() + 1, cols.end());
        Table others;
        decode_sources(ctx, src, rest, &windows, &others, &st);
        for (auto& c : others.cols) t.cols.push_back(std::move(c));
      }
      n_out = (int64_t)oo[nseg];
      Buf<uint64_t> d_oo(ctx, nseg + 1);
      copy_h2d(ctx, d_oo.get(), oo.data(), 8 * (nseg + 1));
      idx.alloc(ctx, std::max<int64_t>(1, n_out));
      if (nseg) {
        k_ranges_to_indices<<<nseg, 256, 0, ctx->stream>>>(d_bounds.get(), d_seg.get(), d_oo.get(), nseg, idx.get());
        HS_LAUNCH_CHECK(ctx);
      }
      sync_stream(ctx);
    } else {
      // full predicate scan (appended source files under Hybrid Scan, or lineage NOT-IN filter)
      Buf<uint32_t> mask(ctx, std::max<int64_t>(1, n));
      Buf<uint64_t> offs(ctx, n + 1);
      if (str_key)
        launch_filter_mask_strings(ctx, (const uint64_t*)t.cols[0].data.get(), t.cols[0].has_nulls ? t.cols[0].valid.get() : nullptr, n,
                                   spec->has_lo, lo_ref, spec->has_hi, hi_ref, mask.get());
      else
        launch_filter_mask(ctx, d_keys, t.cols[0].has_nulls ? t.cols[0].valid.get() : nullptr, n, spec->has_lo, spec->lo,
                           spec->has_hi, spec->hi, mask.get());
      if (spec->n_deleted_file_ids > 0) {
        Buf<int64_t> d_del(ctx, spec->n_deleted_file_ids);
        copy_h2d(ctx, d_del.get(), spec->deleted_file_ids, 8 * spec->n_deleted_file_ids);
        launch_not_in_mask(ctx, (const int64_t*)t.cols[lineage_col].data.get(), n, d_del.get(), spec->n_deleted_file_ids, mask.get());
        sync_stream(ctx);
      }
      exclusive_scan_u32_u64(ctx, mask.get(), n, offs.get());
      uint64_t kept = 0;
      copy_d2h(ctx, &kept, offs.get() + n, 8);
      sync_stream(ctx);
      n_out = (int64_t)kept;
      idx.alloc(ctx, std::max<int64_t>(1, n_out));
      launch_compact_indices(ctx, mask.get(), offs.get(), n, idx.get());
      sync_stream(ctx);
    }
    t_scan.stop();
    batch_from_gather(ctx, t, proj_idx, idx.get(), n_out, res.get());
    total.stop();
    sync_stream(ctx);
    st.ms_sort += t_scan.ms();
    st.rows_out = n_out;
    st.ms_total = total.ms();
    st.gpu_launches = ctx->launches;
  });
  if (stats) *stats = st;
  if (rc == HS_OK) *out = res.release();
  return rc;
}

// Orders the decoded rows of one join side bucket-major and key-sorted.  When every bucket holds exactly one file the
// files are already sorted (they are index files) and only need to be visited in bucket order; otherwise the rows go
// through K2-K4 again, which is what Spark's SortExec does for multi-file buckets.
static void prepare_join_side(hs_ctx* ctx, SourceSet* src, const hs_source_file* files, int n_files, const int32_t* buckets, int nb,
                              const std::vector<std::string>& cols, Table* t, IndexedRows* rows, hs_stats* st,
                              std::vector<uint64_t>* seg, const int64_t** d_keys, const uint32_t** d_perm, Buf<uint32_t>* iota,
                              Buf<int64_t>* k64) {
  // reorder files by bucket so the decoded table is bucket-major
  std::vector<int> order(n_files);
  for (int i = 0; i < n_files; i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return buckets[a] < buckets[b]; });
  std::vector<hs_source_file> sorted_files(n_files);
  std::vector<int> per_bucket(nb, 0);
  for (int i = 0; i < n_files; i++) {
    sorted_files[i] = files[order[i]];
    if (buckets[order[i]] < 0 || buckets[order[i]] >= nb) fail(HS_EINVAL, "bucket id %d out of range", buckets[order[i]]);
    per_bucket[buckets[order[i]]]++;
  }
  // string values are references into the file images: the caller's SourceSet keeps those alive until the result batch exists
  open_sources(ctx, sorted_files.data(), n_files, src, st);
  decode_sources(ctx, *src, cols, nullptr, t, st);
  if (!t->has_strings) src->release_images();
  const bool str_key = t->cols[0].type == HS_TYPE_STRING;
  if (!str_key && t->cols[0].type != HS_TYPE_INT64 && t->cols[0].type != HS_TYPE_INT32)
    fail(HS_EUNSUPPORTED, "bucket join: key column must be int32, int64 or string");
  if (t->cols[0].has_nulls) fail(HS_EUNSUPPORTED, "bucket join: null join keys are not handled yet");
  const bool single = std::all_of(per_bucket.begin(), per_bucket.end(), [](int c) { return c <= 1; });
  if (single) {
    seg->assign(nb + 1, 0);
    int fi = 0;
    for (int b = 0; b < nb; b++) {
      (*seg)[b] = (uint64_t)t->file_row_begin[fi];
      if (per_bucket[b]) fi++;
    }
    (*seg)[nb] = (uint64_t)t->nrows;
    *d_keys = str_key ? (const int64_t*)t->cols[0].data.get() : widened_key(ctx, t->cols[0], t->nrows, k64);
    iota->alloc(ctx, std::max<int64_t>(1, t->nrows));
    launch_iota_u32(ctx, iota->get(), t->nrows);
    *d_perm = iota->get();
  } else {
    index_rows(ctx, *t, 1, nb, rows, st);
    *seg = rows->bucket_offsets;
    // materialise the sorted key column
    const int kw = rows->part.cols[0].width;
    const int ktype = rows->part.cols[0].type;
    Buf<uint8_t> sk(ctx, (size_t)std::max<int64_t>(1, rows->part.nrows) * kw);
    launch_gather_plain(ctx, rows->part.cols[0].data.get(), rows->sorted_perm, rows->part.nrows, kw, sk.get());
    rows->keys_alt.release();
    t->cols.clear();
    t->cols = std::move(rows->part.cols);
    t->nrows = rows->part.nrows;
    // keep the sorted keys in a column appended at the end
    DevColumn kc;
    kc.name = "__sorted_key";
    kc.type = ktype;
    kc.width = kw;
    kc.data = std::move(sk);
    t->cols.push_back(std::move(kc));
    *d_keys = str_key ? (const int64_t*)t->cols.back().data.get() : widened_key(ctx, t->cols.back(), t->nrows, k64);
    *d_perm = rows->sorted_perm;
  }
}

__global__ void k_compose_u32(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, int64_t n, uint32_t* out) {
  // out[i] = a[b[i]]
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = a[b[i]];
}

int hs_bucket_join(hs_ctx* ctx, const hs_join_spec* spec, hs_batch** out, hs_stats* stats, char* err, size_t errlen) {
  if (!ctx || !spec || !out) return HS_EINVAL;
  *out = nullptr;
  hs_stats st;
  memset(&st, 0, sizeof st);
  std::unique_ptr<hs_batch> res(new hs_batch());
  res->ctx = ctx;
  res->on_device = spec->output == HS_OUT_DEVICE;
  int rc = guarded(ctx, err, errlen, [&] {
    StageTimer total(ctx);
    total.start();
    const int nb = spec->num_buckets;
    if (nb < 1) fail(HS_EINVAL, "num_buckets must be positive");
    std::vector<std::string> lcols{spec->left_key}, rcols{spec->right_key};
    std::vector<int> lproj, rproj;
    for (int i = 0; i < spec->n_left_columns; i++) {
      std::string nm = spec->left_columns[i];
      auto it = std::find(lcols.begin(), lcols.end(), nm);
      if (it == lcols.end()) { lcols.push_back(nm); lproj.push_back((int)lcols.size() - 1); }
      else lproj.push_back((int)(it - lcols.begin()));
    }
    for (int i = 0; i < spec->n_right_columns; i++) {
      std::string nm = spec->right_columns[i];
      auto it = std::find(rcols.begin(), rcols.end(), nm);
      if (it == rcols.end()) { rcols.push_back(nm); rproj.push_back((int)rcols.size() - 1); }
      else rproj.push_back((int)(it - rcols.begin()));
    }
    Table lt, rt;
    IndexedRows lrows, rrows;
    std::vector<uint64_t> lseg, rseg;
    const int64_t *lkeys = nullptr, *rkeys = nullptr;
    const uint32_t *lperm = nullptr, *rperm = nullptr;
    Buf<uint32_t> liota, riota;
    Buf<int64_t> lk64, rk64;
    SourceSet lsrc, rsrc;
    prepare_join_side(ctx, &lsrc, spec->left_files, spec->n_left, spec->left_buckets, nb, lcols, &lt, &lrows, &st, &lseg, &lkeys, &lperm, &liota, &lk64);
    prepare_join_side(ctx, &rsrc, spec->right_files, spec->n_right, spec->right_buckets, nb, rcols, &rt, &rrows, &st, &rseg, &rkeys, &rperm, &riota, &rk64);
    // hashInt and hashLong put equal values into different buckets: both sides must have been bucketed on the same type
    // (JoinIndexRule only pairs indexes whose indexed columns have the same data type)
    if (lt.cols[0].type != rt.cols[0].type) fail(HS_EUNSUPPORTED, "bucket join: key columns have different types");
    const int64_t nl = lt.nrows, nr = rt.nrows;
    if (nr >= (1ll << 32) || nl >= (1ll << 32)) fail(HS_EUNSUPPORTED, "join side larger than 2^32-1 rows");
    StageTimer t_join(ctx);
    t_join.start();
    Buf<uint64_t> d_lseg(ctx, nb + 1), d_rseg(ctx, nb + 1);
    copy_h2d(ctx, d_lseg.get(), lseg.data(), 8 * (nb + 1));
    copy_h2d(ctx, d_rseg.get(), rseg.data(), 8 * (nb + 1));
    Buf<uint32_t> counts(ctx, std::max<int64_t>(1, nl)), first(ctx, std::max<int64_t>(1, nl));
    Buf<uint64_t> offs(ctx, nl + 1);
    launch_join_count(ctx, lkeys, d_lseg.get(), rkeys, d_rseg.get(), nb, nl, counts.get(), first.get(),
                      lt.cols[0].type == HS_TYPE_STRING);
    exclusive_scan_u32_u64(ctx, counts.get(), nl, offs.get());
    uint64_t total_out = 0;
    copy_d2h(ctx, &total_out, offs.get() + nl, 8);
    sync_stream(ctx);
    if (total_out >= (1ull << 32)) fail(HS_EUNSUPPORTED, "join output larger than 2^32-1 rows per call");
    Buf<uint32_t> li(ctx, std::max<uint64_t>(1, total_out)), ri(ctx, std::max<uint64_t>(1, total_out));
    launch_join_emit(ctx, counts.get(), first.get(), offs.get(), nl, li.get(), ri.get());
    // positions in sorted order -> rows of the partitioned tables
    Buf<uint32_t> lrow(ctx, std::max<uint64_t>(1, total_out)), rrow(ctx, std::max<uint64_t>(1, total_out));
    if (total_out) {
      const int grid = (int)std::min<int64_t>(ceil_div((int64_t)total_out, 256), ctx->sm_count * 16);
      k_compose_u32<<<grid, 256, 0, ctx->stream>>>(lperm, li.get(), (int64_t)total_out, lrow.get());
      HS_LAUNCH_CHECK(ctx);
      k_compose_u32<<<grid, 256, 0, ctx->stream>>>(rperm, ri.get(), (int64_t)total_out, rrow.get());
      HS_LAUNCH_CHECK(ctx);
    }
    t_join.stop();
    batch_from_gather(ctx, lt, lproj, lrow.get(), (int64_t)total_out, res.get());
    batch_from_gather(ctx, rt, rproj, rrow.get(), (int64_t)total_out, res.get());
    total.stop();
    sync_stream(ctx);
    st.ms_sort += t_join.ms();
    st.rows_out = (int64_t)total_out;
    st.ms_total = total.ms();
    st.gpu_launches = ctx->launches;
  });
  if (stats) *stats = st;
  if (rc == HS_OK) *out = res.release();
  return rc;
}

int64_t hs_batch_num_rows(const hs_batch* b) { return b ? b->nrows : 0; }
int32_t hs_batch_on_device(const hs_batch* b) { return b && b->on_device ? 1 : 0; }
int32_t hs_batch_num_columns(const hs_batch* b) { return b ? (int32_t)b->cols.size() : 0; }
int hs_batch_column(const hs_batch* b, int32_t i, const char** name, int32_t* type, const void** data,
                    const uint8_t** valid) {
  if (!b || i < 0 || i >= (int32_t)b->cols.size()) return HS_EINVAL;
  const hs_batch::Col& c = b->cols[i];
  if (name) *name = c.name.c_str();
  if (type) *type = c.type;
  if (data) *data = c.data.get();
  if (valid) *valid = c.has_valid ? c.valid.get() : nullptr;
  return HS_OK;
}
int hs_batch_string_offsets(const hs_batch* b, int32_t i, const uint64_t** offsets, uint64_t* total_bytes) {
  if (!b || i < 0 || i >= (int32_t)b->cols.size() || b->cols[i].type != HS_TYPE_STRING) return HS_EINVAL;
  if (offsets) *offsets = b->cols[i].offsets.get();
  if (total_bytes) *total_bytes = b->cols[i].total_bytes;
  return HS_OK;
}

void hs_batch_free(hs_batch* b) {
  if (!b) return;
  if (b->ctx) {
    cudaSetDevice(b->ctx->device);
    if (b->on_device) cudaStreamSynchronize(b->ctx->stream);
  }
  delete b;
}

// ---------------------------------------------------------------------------------------------------------------------
// kernel-level entry points
// ---------------------------------------------------------------------------------------------------------------------

static void upload_table(hs_ctx* ctx, const hs_host_column* keys, int nkeys, int64_t nrows, Table* t) {
  t->nrows = nrows;
  t->cols.resize(nkeys);
  for (int k = 0; k < nkeys; k++) {
    DevColumn& c = t->cols[k];
    c.name = "k" + std::to_string(k);
    c.type = keys[k].type;
    c.width = type_width(c.type);
    if (c.width == 0) fail(HS_EUNSUPPORTED, "key type %d is not handled by the GPU path", c.type);
    c.data.alloc(ctx, (size_t)nrows * c.width + 16);
    if (nrows) copy_h2d(ctx, c.data.get(), keys[k].data, (size_t)nrows * c.width);
    if (keys[k].valid) {
      c.valid.alloc(ctx, (size_t)nrows + 16);
      if (nrows) copy_h2d(ctx, c.valid.get(), keys[k].valid, (size_t)nrows);
      c.has_nulls = true;
    }
  }
  sync_stream(ctx);
}

int hs_k_bucket_ids(hs_ctx* ctx, const hs_host_column* keys, int32_t nkeys, int64_t nrows, int32_t num_buckets,
                    int32_t* out_bucket, int64_t* out_hist, char* err, size_t errlen) {
  if (!ctx) return HS_EINVAL;
  return guarded(ctx, err, errlen, [&] {
    if (num_buckets < 1 || num_buckets > kMaxBuckets) fail(HS_EUNSUPPORTED, "numBuckets must be in 1..%d", kMaxBuckets);
    Table t;
    upload_table(ctx, keys, nkeys, nrows, &t);
    std::vector<KeyColumn> h_keys(nkeys);
    for (int k = 0; k < nkeys; k++)
      h_keys[k] = KeyColumn{t.cols[k].data.get(), t.cols[k].has_nulls ? t.cols[k].valid.get() : nullptr, t.cols[k].type, t.cols[k].width};
    Buf<KeyColumn> d_keys(ctx, nkeys);
    copy_h2d(ctx, d_keys.get(), h_keys.data(), sizeof(KeyColumn) * nkeys);
    const int64_t ntiles = ceil_div(nrows, kPartTile);
    Buf<uint16_t> bucket(ctx, std::max<int64_t>(1, nrows));
    Buf<uint32_t> tile_hist(ctx, std::max<int64_t>(1, ntiles) * num_buckets);
    Buf<unsigned long long> ghist(ctx, num_buckets);
    fill_bytes(ctx, ghist.get(), 0, 8 * num_buckets);
    launch_bucket_hist(ctx, d_keys.get(), nkeys, nrows, num_buckets, bucket.get(), tile_hist.get(), ghist.get());
    std::vector<uint16_t> hb(nrows);
    if (nrows) copy_d2h(ctx, hb.data(), bucket.get(), 2 * nrows);
    if (out_hist) copy_d2h(ctx, out_hist, ghist.get(), 8 * num_buckets);
    sync_stream(ctx);
    if (out_bucket) for (int64_t i = 0; i < nrows; i++) out_bucket[i] = hb[i];
  });
}

int hs_k_sort_perm(hs_ctx* ctx, const hs_host_column* keys, int32_t nkeys, int64_t nrows, int32_t num_buckets,
                   int64_t* out_perm, int64_t* out_bucket_offsets, char* err, size_t errlen) {
  if (!ctx) return HS_EINVAL;
  return guarded(ctx, err, errlen, [&] {
    Table t;
    upload_table(ctx, keys, nkeys, nrows, &t);
    // carry the original row index through the partition as an extra column
    DevColumn rid;
    rid.name = "__row";
    rid.type = HS_TYPE_INT32;
    rid.width = 4;
    rid.data.alloc(ctx, (size_t)nrows * 4 + 16);
    launch_iota_u32(ctx, (uint32_t*)rid.data.get(), nrows);
    t.cols.push_back(std::move(rid));
    IndexedRows rows;
    hs_stats st;
    memset(&st, 0, sizeof st);
    index_rows(ctx, t, nkeys, num_buckets, &rows, &st);
    Buf<uint8_t> orig(ctx, (size_t)std::max<int64_t>(1, nrows) * 4);
    launch_gather_plain(ctx, rows.part.cols[nkeys].data.get(), rows.sorted_perm, nrows, 4, orig.get());
    std::vector<uint32_t> h(nrows);
    if (nrows) copy_d2h(ctx, h.data(), orig.get(), 4 * nrows);
    sync_stream(ctx);
    for (int64_t i = 0; i < nrows; i++) out_perm[i] = h[i];
    for (int b = 0; b <= num_buckets; b++) out_bucket_offsets[b] = (int64_t)rows.bucket_offsets[b];
  });
}

int hs_synth_table(hs_ctx* ctx, int64_t first_row, int64_t nrows, int32_t ncols, int32_t n_files,
                   int32_t row_groups_per_file, int32_t dictionary, int32_t output, hs_index_result** out, char* err,
                   size_t errlen) {
  return hs_synth_table_ex(ctx, first_row, nrows, ncols, n_files, row_groups_per_file, dictionary, HS_CODEC_UNCOMPRESSED, output, out,
                           err, errlen);
}

int hs_k_snappy_compress(hs_ctx* ctx, const void* in, uint64_t n, void* out_buf, uint64_t cap, uint64_t* out_len, char* err,
                         size_t errlen) {
  if (!ctx || !out_len || (n && !in)) return HS_EINVAL;
  return guarded(ctx, err, errlen, [&] {
    std::vector<SnappyFragment> frags;
    uint64_t slot = 0;
    for (uint64_t o = 0; o < n; o += kSnappyFragment) {
      const uint32_t len = (uint32_t)std::min<uint64_t>(kSnappyFragment, n - o);
      frags.push_back(SnappyFragment{o, slot, len, 0});
      slot += round_up(snappy_max_compressed(len), 16);
    }
    Buf<uint8_t> d_in(ctx, std::max<uint64_t>(n, 16) + 16), d_slots(ctx, std::max<uint64_t>(slot, 16));
    Buf<SnappyFragment> d_frags(ctx, std::max<size_t>(1, frags.size()));
    Buf<uint32_t> d_len(ctx, std::max<size_t>(1, frags.size()));
    std::vector<uint32_t> lens(frags.size());
    if (n) copy_h2d(ctx, d_in.get(), in, n);
    copy_h2d(ctx, d_frags.get(), frags.data(), sizeof(SnappyFragment) * frags.size());
    launch_snappy_compress(ctx, d_frags.get(), (int64_t)frags.size(), d_in.get(), d_slots.get(), d_len.get());
    copy_d2h(ctx, lens.data(), d_len.get(), 4 * frags.size());
    sync_stream(ctx);
    std::vector<uint8_t> stream;
    for (uint64_t v = n;; v >>= 7) {  // preamble: the uncompressed length
      if (v >= 0x80) stream.push_back((uint8_t)(v | 0x80));
      else {
        stream.push_back((uint8_t)v);
        break;
      }
    }
    std::vector<uint8_t> piece;
    for (size_t f = 0; f < frags.size(); f++) {
      piece.resize(lens[f]);
      HS_CUDA(cudaMemcpy(piece.data(), d_slots.get() + frags[f].dst_off, lens[f], cudaMemcpyDeviceToHost));
      stream.insert(stream.end(), piece.begin(), piece.end());
    }
    *out_len = stream.size();
    if (stream.size() > cap) fail(HS_ENOMEM, "output buffer too small: %zu bytes needed", stream.size());
    if (out_buf) memcpy(out_buf, stream.data(), stream.size());
  });
}

int hs_k_snappy_decompress(hs_ctx* ctx, const void* in, uint64_t n, void* out_buf, uint64_t out_len, int32_t* sequential,
                           char* err, size_t errlen) {
  if (!ctx || !in || n == 0 || n > 0xffffffffull || out_len > 0xfffffff0ull || (out_len && !out_buf)) return HS_EINVAL;
  return guarded(ctx, err, errlen, [&] {
    Buf<uint8_t> d_in(ctx, n + 16), d_out(ctx, out_len + 32);
    copy_h2d(ctx, d_in.get(), in, n);
    SnappyBlob blob{d_in.get(), 0, (uint32_t)n, (uint32_t)out_len, 0u, 1u, 0u, 0u};
    const int64_t blocks = snappy_blocks_of(blob.dst_len, 0);
    Buf<SnappyBlob> d_blob(ctx, 1);
    Buf<uint32_t> d_block_in(ctx, (size_t)blocks + 1), d_seq(ctx, 1), d_error(ctx, 1);
    copy_h2d(ctx, d_blob.get(), &blob, sizeof blob);
    fill_bytes(ctx, d_error.get(), 0, 4);
    launch_snappy_decompress(ctx, d_blob.get(), 1, blocks, false, d_block_in.get(), d_seq.get(), d_out.get(), d_error.get());
    uint32_t error = 0, seq = 0;
    copy_d2h(ctx, &error, d_error.get(), 4);
    copy_d2h(ctx, &seq, d_seq.get(), 4);
    sync_stream(ctx);
    if (error) fail(HS_EFORMAT, "corrupt snappy stream (check %u)", error & 0xffffffu);
    if (out_len) HS_CUDA(cudaMemcpy(out_buf, d_out.get(), out_len, cudaMemcpyDeviceToHost));
    if (sequential) *sequential = (int32_t)seq;
  });
}

int hs_synth_table_ex(hs_ctx* ctx, int64_t first_row, int64_t nrows, int32_t ncols, int32_t n_files,
                      int32_t row_groups_per_file, int32_t dictionary, int32_t compression, int32_t output,
                      hs_index_result** out, char* err, size_t errlen) {
  if (!ctx || !out) return HS_EINVAL;
  *out = nullptr;
  std::unique_ptr<hs_index_result> res(new hs_index_result());
  int rc = guarded(ctx, err, errlen, [&] {
    if (ncols < 1 || ncols > 5 || n_files < 1 || row_groups_per_file < 1 || nrows < 0) fail(HS_EINVAL, "bad synthetic table shape");
    if (output != HS_OUT_HOST && output != HS_OUT_DEVICE) fail(HS_EINVAL, "synthetic tables are returned in memory");
    static const char* names[5] = {"k", "v1", "v2", "v3", "v4"};
    static const int types[5] = {HS_TYPE_INT64, HS_TYPE_INT64, HS_TYPE_DOUBLE, HS_TYPE_INT32, HS_TYPE_FLOAT};
    static const int ptypes[5] = {pq::INT64, pq::INT64, pq::DOUBLE, pq::INT32, pq::FLOAT};
    Table t;
    t.nrows = nrows;
    t.cols.resize(ncols);
    for (int c = 0; c < ncols; c++) {
      DevColumn& dc = t.cols[c];
      dc.name = names[c];
      dc.type = types[c];
      dc.width = type_width(types[c]);
      dc.schema.name = names[c];
      dc.schema.type = ptypes[c];
      dc.schema.repetition = pq::OPTIONAL;
      dc.data.alloc(ctx, (size_t)nrows * dc.width + 16);
      launch_synth_column(ctx, c, first_row, nrows, dc.data.get());
    }
    // files are the segments; rows split evenly, row groups per file likewise (multiples of the page size)
    const int64_t P = 131072;
    std::vector<uint64_t> seg(n_files + 1, 0);
    const int64_t per_file = ceil_div(nrows, n_files);
    for (int f = 0; f < n_files; f++) seg[f + 1] = (uint64_t)std::min<int64_t>(nrows, (int64_t)(f + 1) * per_file);
    SortPlan plan;
    build_sort_plan(ctx, seg.data(), n_files, &plan);
    Buf<uint32_t> iota(ctx, std::max<int64_t>(1, nrows));
    launch_iota_u32(ctx, iota.get(), nrows);
    EncodeRequest req;
    req.table = &t;
    req.d_perm = iota.get();
    req.plan = &plan;
    req.seg_offsets = seg;
    req.rows_per_page = P;
    req.use_dictionary = dictionary != 0;
    if (compression != HS_CODEC_UNCOMPRESSED && compression != HS_CODEC_SNAPPY) fail(HS_EUNSUPPORTED, "compression codec %d", compression);
    req.codec = compression == HS_CODEC_SNAPPY ? pq::SNAPPY : pq::UNCOMPRESSED;
    req.rows_per_row_group = std::max<int64_t>(P, (int64_t)round_up((size_t)ceil_div(per_file, row_groups_per_file), (size_t)P));
    req.seg_names.resize(n_files);
    for (int f = 0; f < n_files; f++) {
      char nm[64];
      snprintf(nm, sizeof nm, "part-%05d.parquet", f);
      req.seg_names[f] = nm;
    }
    EncodedFiles enc;
    hs_stats st;
    memset(&st, 0, sizeof st);
    encode_segments(ctx, req, &enc, &st);
    finish_result(ctx, enc, output, nullptr, HS_SAVE_OVERWRITE, res.get(), &st);
  });
  if (rc == HS_OK) *out = res.release();
  return rc;
}

}  // extern "C"
