// Frame-parallel stream driver (BASELINE config 5): the reference processes one stereo pair per
// StereoMatching object (main_.cpp:138-166); a stream of pairs is embarrassingly parallel BY FRAME -- SGM paths do not
// shard along rows -- so the unit of distribution is the frame, one GPU runs whole frames, and there is no collective.
//
// One worker per device = one host thread + one sm_ctx + one sm_pipeline + a copy stream and two sets of device image
// / result buffers.  Frame i goes to worker i mod n.  Inside a worker the three phases of consecutive frames overlap:
//
//      copy stream :  H2D(i+1)                 D2H(i)          H2D(i+2) ...
//      ctx stream  :            compute(i)            compute(i+1)
//
// H2D(i+1) is enqueued BEFORE compute(i) (it only waits for the compute that last read its buffer set), D2H(i) waits for
// compute(i) through an event, and the host thread blocks only when it retires frame i-1.  Pageable caller buffers go
// through the worker's pinned staging (one host memcpy), pinned ones are used directly.
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

struct job_t {
  const uint8_t *bgrL, *bgrR, *grayL, *grayR;
  int16_t* dispL;
  long long id;
};

struct worker_t {
  int device = 0, H = 0, W = 0;
  sm_params p;
  sm_ctx* ctx = nullptr;
  sm_pipeline* pl = nullptr;
  cudaStream_t copy = nullptr;
  uint8_t* d_in[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
  int16_t* d_out[2] = {nullptr, nullptr};
  uint8_t* h_in[2] = {nullptr, nullptr};
  int16_t* h_out[2] = {nullptr, nullptr};
  cudaEvent_t in_ready[2], computed[2], out_done[2];
  bool in_used[2] = {false, false}, ev_ok = false;
  std::thread th;
  std::mutex m;
  std::condition_variable cv_job, cv_done;
  std::deque<job_t> q;
  bool stop = false;
  int depth = 4;
  long long frames_done = 0;       // retired frames (ids are per-stream; a worker retires its own in order)
  long long last_done_id = -1;
  int err = SM_OK;
  std::string errmsg;
  double busy_ms = 0.0;            // device time between the first enqueue and the last retire (wall clock)
};

bool host_is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

#define W_CUDA(call)                                                                             \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess) {                                                                    \
      sm_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));       \
      return SM_ERR_CUDA;                                                                        \
    }                                                                                            \
  } while (0)

// H2D of one frame into buffer set `s` on the copy stream
int w_prefetch(worker_t* w, int s, const job_t& j) {
  const size_t npix = (size_t)w->H * w->W;
  const size_t sz[4] = {npix * 3, npix * 3, npix, npix};
  const uint8_t* src[4] = {j.bgrL, j.bgrR, j.grayL, j.grayR};
  // the set was last read by the compute two frames ago; its staging copy by that frame's H2D
  if (w->in_used[s]) {
    W_CUDA(cudaStreamWaitEvent(w->copy, w->computed[s], 0));
    W_CUDA(cudaEventSynchronize(w->in_ready[s]));
  }
  size_t off = 0;
  for (int k = 0; k < 4; k++) {
    if (!src[k]) continue;
    const uint8_t* from = src[k];
    if (!host_is_pinned(from)) { memcpy(w->h_in[s] + off, from, sz[k]); from = w->h_in[s] + off; }
    W_CUDA(cudaMemcpyAsync(w->d_in[s][k], from, sz[k], cudaMemcpyHostToDevice, w->copy));
    off += sz[k];
  }
  W_CUDA(cudaEventRecord(w->in_ready[s], w->copy));
  w->in_used[s] = true;
  return SM_OK;
}

// compute of the frame in buffer set `s`, then its D2H on the copy stream
int w_run(worker_t* w, int s, const job_t& j, bool* staged_out) {
  cudaStream_t cs = (cudaStream_t)sm_ctx_stream(w->ctx);
  const size_t npix = (size_t)w->H * w->W;
  W_CUDA(cudaStreamWaitEvent(cs, w->in_ready[s], 0));
  W_CUDA(cudaStreamWaitEvent(cs, w->out_done[s], 0));   // d_out[s] was last read by the D2H two frames ago
  SM_TRY(sm_pipeline_bind_inputs(w->pl, w->d_in[s][0], w->d_in[s][1], j.grayL ? w->d_in[s][2] : nullptr,
                                 j.grayL ? w->d_in[s][3] : nullptr));
  SM_TRY(sm_pipeline_run_device(w->pl));
  W_CUDA(cudaMemcpyAsync(w->d_out[s], sm_pipeline_buffer(w->pl, 2), npix * 2, cudaMemcpyDeviceToDevice, cs));
  W_CUDA(cudaEventRecord(w->computed[s], cs));
  W_CUDA(cudaStreamWaitEvent(w->copy, w->computed[s], 0));
  *staged_out = !host_is_pinned(j.dispL);
  W_CUDA(cudaMemcpyAsync(*staged_out ? w->h_out[s] : j.dispL, w->d_out[s], npix * 2, cudaMemcpyDeviceToHost, w->copy));
  W_CUDA(cudaEventRecord(w->out_done[s], w->copy));
  return SM_OK;
}

int w_retire(worker_t* w, int s, const job_t& j, bool staged_out) {
  W_CUDA(cudaEventSynchronize(w->out_done[s]));
  if (staged_out) memcpy(j.dispL, w->h_out[s], (size_t)w->H * w->W * 2);
  return SM_OK;
}

void w_fail(worker_t* w, int rc) {
  std::lock_guard<std::mutex> lk(w->m);
  if (w->err == SM_OK) { w->err = rc; w->errmsg = sm_last_error(); }
}

void w_mark_done(worker_t* w, const job_t& j) {
  {
    std::lock_guard<std::mutex> lk(w->m);
    w->frames_done++;
    w->last_done_id = j.id;
  }
  w->cv_done.notify_all();
}

void worker_main(worker_t* w) {
  cudaSetDevice(w->device);
  bool have_cur = false, have_prev = false, prev_staged = false;
  job_t cur{}, prev{};
  int s = 0, prev_s = 0;
  for (;;) {
    if (!have_cur) {
      std::unique_lock<std::mutex> lk(w->m);
      w->cv_job.wait(lk, [&] { return w->stop || !w->q.empty(); });
      if (w->q.empty()) break;   // stop requested and nothing left
      cur = w->q.front();
      w->q.pop_front();
      lk.unlock();
      w->cv_done.notify_all();   // a queue slot is free (submit's back-pressure)
      have_cur = true;
      int rc = w_prefetch(w, s, cur);
      if (rc != SM_OK) { w_fail(w, rc); w_mark_done(w, cur); have_cur = false; continue; }
    }
    // the next frame, if one is already waiting: its upload goes out before this frame's compute is enqueued
    job_t next{};
    bool have_next = false;
    {
      std::lock_guard<std::mutex> lk(w->m);
      if (!w->q.empty()) { next = w->q.front(); w->q.pop_front(); have_next = true; }
    }
    int rc = SM_OK;
    if (have_next) { w->cv_done.notify_all(); rc = w_prefetch(w, s ^ 1, next); }
    bool staged = false;
    if (rc == SM_OK) rc = w_run(w, s, cur, &staged);
    if (have_prev) {
      int r2 = w_retire(w, prev_s, prev, prev_staged);
      if (r2 != SM_OK) w_fail(w, r2);
      w_mark_done(w, prev);
      have_prev = false;
    }
    if (rc != SM_OK) {
      w_fail(w, rc);
      cudaStreamSynchronize((cudaStream_t)sm_ctx_stream(w->ctx));
      cudaStreamSynchronize(w->copy);
      w_mark_done(w, cur);
      if (have_next) w_mark_done(w, next);
      have_cur = false;
      continue;
    }
    prev = cur; prev_s = s; prev_staged = staged; have_prev = true;
    if (have_next) { cur = next; s ^= 1; }
    else {
      // nothing queued behind this frame: retire it now so a waiting caller gets it without another submit
      int r2 = w_retire(w, prev_s, prev, prev_staged);
      if (r2 != SM_OK) w_fail(w, r2);
      w_mark_done(w, prev);
      have_prev = false;
      have_cur = false;
      s ^= 1;
    }
  }
}

void worker_free(worker_t* w) {
  if (!w) return;
  cudaSetDevice(w->device);
  if (w->pl) sm_pipeline_destroy(w->pl);
  for (int s = 0; s < 2; s++) {
    for (int k = 0; k < 4; k++) cudaFree(w->d_in[s][k]);
    cudaFree(w->d_out[s]);
    if (w->h_in[s]) cudaFreeHost(w->h_in[s]);
    if (w->h_out[s]) cudaFreeHost(w->h_out[s]);
    if (w->ev_ok) { cudaEventDestroy(w->in_ready[s]); cudaEventDestroy(w->computed[s]); cudaEventDestroy(w->out_done[s]); }
  }
  if (w->copy) cudaStreamDestroy(w->copy);
  if (w->ctx) sm_ctx_destroy(w->ctx);
  delete w;
}

int worker_init(worker_t* w) {
  const size_t npix = (size_t)w->H * w->W;
  SM_TRY(sm_ctx_create(&w->ctx, w->device, nullptr));
  SM_TRY(sm_pipeline_create(w->ctx, w->H, w->W, &w->p, &w->pl));
  W_CUDA(cudaSetDevice(w->device));
  W_CUDA(cudaStreamCreateWithFlags(&w->copy, cudaStreamNonBlocking));
  for (int s = 0; s < 2; s++) {
    const size_t sz[4] = {npix * 3, npix * 3, npix, npix};
    for (int k = 0; k < 4; k++) W_CUDA(cudaMalloc((void**)&w->d_in[s][k], sz[k]));
    W_CUDA(cudaMalloc((void**)&w->d_out[s], npix * 2));
    W_CUDA(cudaMallocHost((void**)&w->h_in[s], npix * 8));
    W_CUDA(cudaMallocHost((void**)&w->h_out[s], npix * 2));
    W_CUDA(cudaEventCreateWithFlags(&w->in_ready[s], cudaEventDisableTiming));
    W_CUDA(cudaEventCreateWithFlags(&w->computed[s], cudaEventDisableTiming));
    W_CUDA(cudaEventCreateWithFlags(&w->out_done[s], cudaEventDisableTiming));
  }
  w->ev_ok = true;
  for (int s = 0; s < 2; s++) W_CUDA(cudaEventRecord(w->out_done[s], w->copy));   // "nothing pending" for the first waits
  return SM_OK;
}

}  // namespace

struct sm_stream {
  std::vector<worker_t*> workers;
  long long next_id = 0;
  int H = 0, W = 0;
};

extern "C" int sm_stream_destroy(sm_stream* s) {
  if (!s) return SM_OK;
  for (worker_t* w : s->workers) {
    if (w->th.joinable()) {
      { std::lock_guard<std::mutex> lk(w->m); w->stop = true; }
      w->cv_job.notify_all();
      w->th.join();
    }
    worker_free(w);
  }
  delete s;
  return SM_OK;
}

extern "C" int sm_stream_create(const int* devices, int n_devices, int H, int W, const sm_params* p, int queue_depth,
                                sm_stream** out) {
  SM_CHECK_ARG(devices && n_devices >= 1 && n_devices <= 64 && p && out && H > 0 && W > 0);
  const int ndev = sm_device_count();
  if (ndev <= 0) { sm_set_error("sm_stream_create: no CUDA device visible (this library has no CPU fallback)"); return SM_ERR_CUDA; }
  for (int i = 0; i < n_devices; i++) SM_CHECK_ARG(devices[i] >= 0 && devices[i] < ndev);
  sm_stream* s = new sm_stream();
  s->H = H; s->W = W;
  for (int i = 0; i < n_devices; i++) {
    worker_t* w = new worker_t();
    w->device = devices[i]; w->H = H; w->W = W; w->p = *p; w->depth = queue_depth > 0 ? queue_depth : 4;
    s->workers.push_back(w);
    int rc = worker_init(w);
    if (rc != SM_OK) { sm_stream_destroy(s); return rc; }
  }
  for (worker_t* w : s->workers) w->th = std::thread(worker_main, w);
  *out = s;
  return SM_OK;
}

extern "C" int sm_stream_submit(sm_stream* s, const uint8_t* h_bgrL, const uint8_t* h_bgrR, const uint8_t* h_grayL,
                                const uint8_t* h_grayR, int16_t* h_dispL, long long* ticket) {
  SM_CHECK_ARG(s && h_bgrL && h_bgrR && h_dispL);
  SM_CHECK_ARG((h_grayL == nullptr) == (h_grayR == nullptr));
  const long long id = s->next_id++;
  worker_t* w = s->workers[(size_t)(id % (long long)s->workers.size())];
  {
    std::unique_lock<std::mutex> lk(w->m);
    w->cv_done.wait(lk, [&] { return (int)w->q.size() < w->depth; });
    w->q.push_back(job_t{h_bgrL, h_bgrR, h_grayL, h_grayR, h_dispL, id});
  }
  w->cv_job.notify_one();
  if (ticket) *ticket = id;
  return SM_OK;
}

extern "C" int sm_stream_wait(sm_stream* s, long long ticket) {
  SM_CHECK_ARG(s && ticket >= 0 && ticket < s->next_id);
  worker_t* w = s->workers[(size_t)(ticket % (long long)s->workers.size())];
  std::unique_lock<std::mutex> lk(w->m);
  w->cv_done.wait(lk, [&] { return w->last_done_id >= ticket; });   // a worker retires its frames in ticket order
  if (w->err != SM_OK) { sm_set_error("sm_stream worker on device %d: %s", w->device, w->errmsg.c_str()); return w->err; }
  return SM_OK;
}

extern "C" int sm_stream_drain(sm_stream* s) {
  SM_CHECK_ARG(s);
  int rc = SM_OK;
  const long long n = s->next_id, nw = (long long)s->workers.size();
  for (long long k = 0; k < nw && k < n; k++) {
    const long long last = n - 1 - ((n - 1 - k) % nw + nw) % nw;   // the last ticket of worker k
    if (last < 0) continue;
    int r = sm_stream_wait(s, last);
    if (r != SM_OK) rc = r;
  }
  return rc;
}

extern "C" int sm_stream_device_count(sm_stream* s) { return s ? (int)s->workers.size() : 0; }

extern "C" long long sm_stream_frames_done(sm_stream* s, int worker) {
  if (!s || worker < 0 || worker >= (int)s->workers.size()) return -1;
  worker_t* w = s->workers[worker];
  std::lock_guard<std::mutex> lk(w->m);
  return w->frames_done;
}

extern "C" long long sm_stream_launch_count(sm_stream* s) {
  if (!s) return 0;
  long long n = 0;
  for (worker_t* w : s->workers) n += sm_ctx_launch_count(w->ctx);
  return n;
}
