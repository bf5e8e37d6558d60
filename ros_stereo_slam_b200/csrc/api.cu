// api.cu -- the C ABI of include/vo_b200.h: context, host-side orchestration of the stages
// (the reference's method boundaries, include/visualSLAM.h:152-169 of the reference tree),
// OpenCV-compatible RANSAC sample generation, and the device-resident sequence driver.
// No CPU compute path exists: every stage launches the CUDA kernels of this library.
#include <float.h>
#include <math.h>
#include <stdarg.h>

#include <cuda.h>   // types of the green-context driver API (entry points come through the runtime)

#include <algorithm>

#include "common.cuh"
#include "cvmath.cuh"

using namespace vo;

// ------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";

namespace vo {
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

LaunchScope::LaunchScope(vo_ctx* ctx, int k) : c(ctx), kind(k) {
  c->launch_count++;
  if (c->prof.mask & (1u << k)) {
    auto get = [&]() {
      cudaEvent_t e;
      if (!c->prof.pool.empty()) {
        e = c->prof.pool.back();
        c->prof.pool.pop_back();
      } else {
        cudaEventCreate(&e);
      }
      return e;
    };
    a = get();
    b = get();
    cudaEventRecord(a, c->stream);
  }
}
LaunchScope::~LaunchScope() {
  if (a) {
    cudaEventRecord(b, c->stream);
    c->prof.pending.push_back({kind, a, b});
  }
}
}  // namespace vo

static void prof_drain(vo_ctx* c) {
  for (auto& p : c->prof.pending) {
    float ms = 0;
    cudaEventSynchronize(p.b);
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      c->prof.launches[p.kind]++;
      c->prof.ms[p.kind] += ms;
    }
    c->prof.pool.push_back(p.a);
    c->prof.pool.push_back(p.b);
  }
  c->prof.pending.clear();
}

static void make_projections(const vo_params& p, double* P /*24*/);
constexpr int F_CHUNK = 48, PNP_CHUNK = 32, F_CHUNK_TEMPORAL = 96;   // first-chunk sizes of the fused chains (see below)

// ------------------------------------------------------------------------------------ SM partition (green contexts)
// VO_B200_ISLAND=<SMs>: the device's SMs are split into an island and the rest; streams created on the two green
// contexts are confined to their SMs.  The driver entry points come through the runtime (no libcuda link).
static int partition_create(vo_ctx* c, int island_sms) {
  typedef CUresult (*fn_devget)(CUdevice*, int);
  typedef CUresult (*fn_getres)(CUdevice, CUdevResource*, CUdevResourceType);
  typedef CUresult (*fn_split)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
  typedef CUresult (*fn_desc)(CUdevResourceDesc*, CUdevResource*, unsigned int);
  typedef CUresult (*fn_gcreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
  typedef CUresult (*fn_gstream)(CUstream*, CUgreenCtx, unsigned int, int);
  void* f[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  const char* names[6] = {"cuDeviceGet", "cuDeviceGetDevResource", "cuDevSmResourceSplitByCount", "cuDevResourceGenerateDesc",
                          "cuGreenCtxCreate", "cuGreenCtxStreamCreate"};
  for (int i = 0; i < 6; i++) {
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(names[i], &f[i], cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !f[i]) {
      cudaGetLastError();
      set_error("SM partition: %s is not available in this driver", names[i]);
      return VO_ERR_CUDA;
    }
  }
  CUdevice dev;
  CUdevResource all, island, rest;
  unsigned int nb = 1;
  CUdevResourceDesc d_island, d_rest;
  CUgreenCtx g_island = nullptr, g_big = nullptr;
#define VO_CU(call)                                                   \
  do {                                                                \
    const CUresult r_ = (call);                                       \
    if (r_ != CUDA_SUCCESS) {                                         \
      set_error("SM partition: %s failed (%d)", #call, (int)r_);      \
      return VO_ERR_CUDA;                                             \
    }                                                                 \
  } while (0)
  VO_CU(((fn_devget)f[0])(&dev, c->device));
  VO_CU(((fn_getres)f[1])(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
  VO_CU(((fn_split)f[2])(&island, &nb, &all, &rest, 0, (unsigned)island_sms));
  VO_CU(((fn_desc)f[3])(&d_island, &island, 1));
  VO_CU(((fn_desc)f[3])(&d_rest, &rest, 1));
  VO_CU(((fn_gcreate)f[4])(&g_island, d_island, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  VO_CU(((fn_gcreate)f[4])(&g_big, d_rest, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  int lo = 0, hi = 0;
  VO_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CUstream si = nullptr, sl = nullptr, sa = nullptr;
  VO_CU(((fn_gstream)f[5])(&si, g_island, CU_STREAM_NON_BLOCKING, hi));
  VO_CU(((fn_gstream)f[5])(&sl, g_big, CU_STREAM_NON_BLOCKING, hi));
  VO_CU(((fn_gstream)f[5])(&sa, g_big, CU_STREAM_NON_BLOCKING, lo));
#undef VO_CU
  c->g_island = g_island;
  c->g_big = g_big;
  c->s_island = (cudaStream_t)si;
  c->s_lk = (cudaStream_t)sl;
  c->s_lk_aux = (cudaStream_t)sa;
  for (int i = 0; i < 4; i++) VO_CUDA(cudaEventCreateWithFlags(&c->ev_p[i], cudaEventDisableTiming));
  c->island_sms = (int)island.sm.smCount;
  if (getenv("VO_B200_DEBUG_FALLBACK"))
    fprintf(stderr, "[vo partition] island %u SMs, rest %u SMs\n", island.sm.smCount, rest.sm.smCount);
  return VO_OK;
}

static void partition_destroy(vo_ctx* c) {
  if (!c->g_island && !c->g_big) return;
  for (cudaStream_t st : {c->s_island, c->s_lk, c->s_lk_aux})
    if (st) {
      cudaStreamSynchronize(st);
      cudaStreamDestroy(st);
    }
  for (int i = 0; i < 4; i++)
    if (c->ev_p[i]) {
      cudaEventDestroy(c->ev_p[i]);
      c->ev_p[i] = nullptr;
    }
  typedef CUresult (*fn_gdestroy)(CUgreenCtx);
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuGreenCtxDestroy", &f, cudaEnableDefault, &q) == cudaSuccess && f) {
    if (c->g_island) ((fn_gdestroy)f)((CUgreenCtx)c->g_island);
    if (c->g_big) ((fn_gdestroy)f)((CUgreenCtx)c->g_big);
  } else {
    cudaGetLastError();
  }
  c->g_island = c->g_big = nullptr;
  c->s_island = c->s_lk = c->s_lk_aux = nullptr;
}

// run `body` with c->stream replaced by `s` (a green-context stream): everything enqueued so far on c->stream precedes it,
// everything enqueued afterwards on c->stream follows it
template <class F>
static int on_stream(vo_ctx* c, cudaStream_t s, cudaEvent_t e_in, cudaEvent_t e_out, F body) {
  VO_CUDA(cudaEventRecord(e_in, c->stream));
  VO_CUDA(cudaStreamWaitEvent(s, e_in, 0));
  cudaStream_t keep = c->stream;
  c->stream = s;
  const int r = body();
  c->stream = keep;
  VO_TRY(r);
  VO_CUDA(cudaEventRecord(e_out, s));
  VO_CUDA(cudaStreamWaitEvent(c->stream, e_out, 0));
  return VO_OK;
}

static int alloc_chain(vo_ctx* c) {
  const vo_params* p = &c->p;
  {
    // The primary chain carries the frame's critical path (tracking -> F -> PnP -> refine), made of
    // short latency-bound kernels; the auxiliary chain carries the stereo LK that fills every SM.
    // Without priorities the critical kernels queue behind the 4,570 CTAs of that LK launch.
    int lo = 0, hi = 0;
    VO_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    VO_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, c->is_aux ? lo : hi));
  }
  const int cap = p->max_points;
  c->cap = cap;
  VO_CUDA(cudaMalloc(&c->d_xy_in, cap * sizeof(float2)));
  VO_CUDA(cudaMalloc(&c->d_xy_trk, cap * sizeof(float2)));
  VO_CUDA(cudaMalloc(&c->d_status, cap));
  VO_CUDA(cudaMalloc(&c->d_err, cap * sizeof(float)));
  VO_CUDA(cudaMalloc(&c->d_xyz_in, cap * sizeof(float3)));
  VO_CUDA(cudaMalloc(&c->d_c_ref, cap * sizeof(float2)));
  VO_CUDA(cudaMalloc(&c->d_c_trk, cap * sizeof(float2)));
  VO_CUDA(cudaMalloc(&c->d_c_xyz, cap * sizeof(float3)));
  VO_CUDA(cudaMalloc(&c->d_f_ref, cap * sizeof(float2)));
  VO_CUDA(cudaMalloc(&c->d_f_trk, cap * sizeof(float2)));
  VO_CUDA(cudaMalloc(&c->d_f_xyz, cap * sizeof(float3)));
  VO_CUDA(cudaMalloc(&c->d_xyz_tmp, cap * sizeof(float3)));
  VO_CUDA(cudaMalloc(&c->d_mask, cap));
  VO_TRY(refine_init());
  VO_CUDA(cudaMalloc(&c->d_idx, cap * sizeof(int32_t)));
  if (!c->is_aux) {
    VO_CUDA(cudaMalloc(&c->d_seq_xy, cap * sizeof(float2)));
    VO_CUDA(cudaMalloc(&c->d_seq_xyz, cap * sizeof(float3)));
  }
  // the small results a fused chain reads back -- pose (16 doubles) | counts (16 ints) | selection (8) | flags (8) -- live
  // in ONE block on either side, so that a chain ends with one device-to-host copy instead of four
  VO_CUDA(cudaMalloc(&c->d_res, RES_BYTES));
  VO_CUDA(cudaMemset(c->d_res, 0, RES_BYTES));
  VO_CUDA(cudaMallocHost(&c->h_res, RES_BYTES));
  memset(c->h_res, 0, RES_BYTES);
  c->d_pose = reinterpret_cast<double*>(c->d_res);
  c->h_pose = reinterpret_cast<double*>(c->h_res);
  c->d_count = reinterpret_cast<int*>(c->d_res + 16 * sizeof(double));
  c->h_count = reinterpret_cast<int*>(c->h_res + 16 * sizeof(double));
  c->d_sel = c->d_count + 16;
  c->h_sel = c->h_count + 16;
  c->d_flags = c->d_count + 24;
  c->h_flags = c->h_count + 24;
  const int n_tiles = std::max(256, (cap + 2047) / 2048 + 1);   // compaction tiles of 2048 flags (points.cu CP_TILE)
  VO_CUDA(cudaMalloc(&c->d_tile_state, n_tiles * sizeof(unsigned long long)));
  {
    const unsigned one[4] = {1, 0, 0, 0};
    VO_CUDA(cudaMalloc(&c->d_epoch, sizeof(one)));
    VO_CUDA(cudaMemcpy(c->d_epoch, one, sizeof(one), cudaMemcpyHostToDevice));
    double P[24];
    make_projections(c->p, P);
    VO_CUDA(cudaMalloc(&c->d_Pst, sizeof(P)));
    VO_CUDA(cudaMemcpy(c->d_Pst, P, sizeof(P), cudaMemcpyHostToDevice));
  }
  VO_CUDA(cudaMemsetAsync(c->d_tile_state, 0, n_tiles * sizeof(unsigned long long), c->stream));
  VO_CUDA(cudaMallocHost(&c->h_pts, (size_t)cap * 8 * sizeof(float)));
  const int ch = p->max_hypotheses;
  c->cap_h = ch;
  VO_CUDA(cudaMalloc(&c->d_samples, (size_t)ch * 7 * sizeof(int32_t)));
  VO_CUDA(cudaMemset(c->d_samples, 0, (size_t)ch * 7 * sizeof(int32_t)));   // never a stale index in front of a gather
  VO_CUDA(cudaMallocHost(&c->h_samples, (size_t)ch * 7 * sizeof(int32_t)));
  VO_CUDA(cudaMalloc(&c->d_models, (size_t)ch * 27 * sizeof(double)));
  VO_CUDA(cudaMalloc(&c->d_counts, (size_t)ch * 3 * sizeof(int32_t)));

  VO_CUDA(cudaMalloc(&c->d_cam, 48 * sizeof(double)));
  {
    std::vector<uint32_t> raw(RNG_LEN);
    CvRng rng(0xffffffffffffffffULL);
    for (int i = 0; i < RNG_LEN; i++) raw[i] = rng.next();
    VO_CUDA(cudaMalloc(&c->d_rng, RNG_LEN * sizeof(uint32_t)));
    VO_CUDA(cudaMemcpy(c->d_rng, raw.data(), RNG_LEN * sizeof(uint32_t), cudaMemcpyHostToDevice));

  }
  if (p->channels == 3) VO_CUDA(cudaMalloc(&c->d_bgr, (size_t)3 * p->width * p->height));
  VO_CUDA(cudaMalloc(&c->d_lk_work, 4 * sizeof(unsigned long long)));
  VO_CUDA(cudaMallocHost(&c->h_lk_work, 4 * sizeof(unsigned long long)));
  VO_CUDA(cudaMemsetAsync(c->d_lk_work, 0, 4 * sizeof(unsigned long long), c->stream));
  VO_CUDA(cudaStreamSynchronize(c->stream));
  return VO_OK;
}

static void free_chain(vo_ctx* c) {
  if (!c) return;
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (int i = 0; i < 4; i++)      // partition events of an auxiliary chain (the streams belong to the main context)
    if (c->ev_p[i]) {
      cudaEventDestroy(c->ev_p[i]);
      c->ev_p[i] = nullptr;
    }
  sgbm_free(c);
  orb_free(c);
  void* dev[] = {c->d_xy_in, c->d_xy_trk, c->d_status, c->d_err, c->d_xyz_in, c->d_c_ref, c->d_c_trk, c->d_c_xyz,
                 c->d_f_ref, c->d_f_trk, c->d_f_xyz, c->d_xyz_tmp, c->d_mask, c->d_idx, c->d_seq_xy, c->d_seq_xyz,
                 c->d_res, c->d_tile_state, c->d_samples, c->d_models, c->d_counts, c->d_cam,
                 c->d_lk_work, c->d_rng, c->d_epoch, c->d_Pst, c->d_bgr, c->d_gray, c->d_sor};
  for (void* p : dev) cudaFree(p);
  void* host[] = {c->h_res, c->h_pts, c->h_samples, c->h_lk_work};
  for (void* p : host) cudaFreeHost(p);
  for (auto& pe : c->prof.pending) {
    cudaEventDestroy(pe.a);
    cudaEventDestroy(pe.b);
  }
  for (auto e : c->prof.pool) cudaEventDestroy(e);
  if (c->stream) cudaStreamDestroy(c->stream);
}

// host worker of the auxiliary chain: runs one posted task at a time
static void worker_main(vo_ctx* c) {
  cudaSetDevice(c->device);
  std::unique_lock<std::mutex> lk(c->mtx);
  for (;;) {
    c->cv.wait(lk, [&] { return c->task_pending || c->quit; });
    if (c->quit) return;
    std::function<int()> f = c->task;
    c->task_pending = false;
    lk.unlock();
    const int r = f();
    const std::string err = r != VO_OK ? std::string(g_err) : std::string();
    lk.lock();
    c->task_result = r;
    c->task_error = err;
    c->task_done = true;
    c->cv.notify_all();
  }
}

static void post_task(vo_ctx* c, std::function<int()> f) {
  std::lock_guard<std::mutex> g(c->mtx);
  c->task = std::move(f);
  c->task_pending = true;
  c->task_done = false;
  c->cv.notify_all();
}

static int wait_task(vo_ctx* c) {
  std::unique_lock<std::mutex> lk(c->mtx);
  c->cv.wait(lk, [&] { return c->task_done; });
  c->task_done = false;
  if (c->task_result != VO_OK) set_error("%s", c->task_error.c_str());
  return c->task_result;
}

extern "C" {

const char* vo_last_error(void) { return g_err; }
int vo_abi_version(void) { return VO_B200_ABI_VERSION; }

const char* vo_strerror(int code) {
  switch (code) {
    case VO_OK: return "ok";
    case VO_ERR_INVALID_ARG: return "invalid argument";
    case VO_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU path)";
    case VO_ERR_CUDA: return "CUDA error";
    case VO_ERR_CAPACITY: return "capacity exceeded";
    case VO_ERR_TOO_FEW_POINTS: return "too few points";
    case VO_ERR_NO_MODEL: return "RANSAC found no model";
    case VO_ERR_LOW_INLIERS: return "low inlier count after retry (reference SHUTDOWN_FLAG)";
    case VO_ERR_NOT_IMPLEMENTED: return "not implemented";
    case VO_ERR_SELF_CHECK: return "numerics self-check failed";
  }
  return "unknown";
}

void vo_default_params(vo_params* p) {
  memset(p, 0, sizeof(*p));
  p->fx = 7.188560000000e+02;
  p->fy = 7.188560000000e+02;
  p->cx = 6.071928000000e+02;
  p->cy = 1.852157000000e+02;
  p->baseline = 0.54;
  p->width = 1241;
  p->height = 376;
  p->channels = 1;
  p->lk_win = 21;
  p->lk_max_level = 3;
  p->lk_max_iters = 30;
  p->lk_eps = 0.01;
  p->lk_min_eig = 1e-4;
  p->grid_step = 30;
  p->f_thr_stereo = 3.0;
  p->f_thr_temporal = 1.0;
  p->f_conf = 0.99;
  p->f_max_iters = 1000;
  p->pnp_iters = 100;
  p->pnp_thr = 1.0;
  p->pnp_conf = 0.99;
  p->pnp_retry_iters = 100;
  p->pnp_retry_thr = 8.0;
  p->pnp_retry_conf = 0.98;
  p->pnp_min_inliers = 10;
  p->kf_min_inliers = 200;
  p->ransac_exhaustive = 0;
  p->f_exhaustive = 0;
  p->max_points = 131072;
  p->max_hypotheses = 4096;
  p->device = 0;
}

// ------------------------------------------------------------------------------------ ctx
int vo_create(const vo_params* p, vo_ctx** out) {
  if (!p || !out) return VO_ERR_INVALID_ARG;
  *out = nullptr;
  if (p->channels != 1 && p->channels != 3) {
    set_error("channels=%d: images are 1-channel (gray) or 3-channel (BGR, what the reference's imread returns)",
              p->channels);
    return VO_ERR_INVALID_ARG;
  }
  if (p->lk_win != LK_WIN) {
    set_error("lk_win=%d: the LK kernel is specialised for the reference's 21x21 window", p->lk_win);
    return VO_ERR_NOT_IMPLEMENTED;
  }
  if (p->width < 64 || p->height < 64 || p->max_points < 32 || p->max_hypotheses < 1 || p->lk_max_level < 0) {
    set_error("bad geometry/capacity");
    return VO_ERR_INVALID_ARG;
  }
  if (p->lk_max_level > MAX_LEVELS - 1) {
    set_error("lk_max_level=%d: the pyramid holds at most %d levels (the reference uses maxLevel 3)", p->lk_max_level, MAX_LEVELS);
    return VO_ERR_NOT_IMPLEMENTED;
  }
  {
    // the single-synchronisation chains write their first chunk of hypotheses without asking the host: the
    // buffers must hold it (the host-driven loops check their own sizes per call)
    const int f_first = p->f_exhaustive ? std::min(std::max(p->f_max_iters, 1), 1024) : std::min(F_CHUNK_TEMPORAL, std::max(p->f_max_iters, 1));
    const int it = std::max(p->pnp_iters, 1);
    const int p_first = p->ransac_exhaustive ? (it <= 1024 ? it : 0) : std::min(it, PNP_CHUNK);
    const int need = std::max(f_first, p_first);
    if (p->max_hypotheses < need) {
      set_error("max_hypotheses=%d is below the %d hypotheses the fused chains evaluate in their first chunk", p->max_hypotheses, need);
      return VO_ERR_INVALID_ARG;
    }
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    set_error("no CUDA device visible; libvo_b200 has no CPU fallback");
    return VO_ERR_NO_DEVICE;
  }
  if (p->device < 0 || p->device >= ndev) {
    set_error("device %d out of range (%d devices)", p->device, ndev);
    return VO_ERR_INVALID_ARG;
  }
  VO_CUDA(cudaSetDevice(p->device));
  vo_ctx* c = new vo_ctx();
  c->p = *p;
  c->device = p->device;
  cudaDeviceProp prop;
  VO_CUDA(cudaGetDeviceProperties(&prop, p->device));
  c->sm_count = prop.multiProcessorCount;
  c->pyr = new Pyramid[4];
  int r = alloc_chain(c);
  if (r == VO_OK) {
    for (int s = 0; s < 4 && r == VO_OK; s++) r = pyr_alloc(c, c->pyr[s]);
  }
  if (r == VO_OK) {
    vo_ctx* a = new vo_ctx();
    a->p = *p;
    a->device = p->device;
    a->sm_count = c->sm_count;
    a->is_aux = true;
    a->pyr = c->pyr;
    c->aux = a;
    r = alloc_chain(a);
  }
  if (r == VO_OK) {
    vo_ctx* l = new vo_ctx();     // look-ahead chain (see common.cuh)
    l->p = *p;
    l->device = p->device;
    l->sm_count = c->sm_count;
    l->is_aux = true;
    l->pyr = c->pyr;
    c->la = l;
    r = alloc_chain(l);
  }
  if (r == VO_OK) {
    if (cudaEventCreateWithFlags(&c->ev_left, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_lk, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_xform, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_la, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_gather, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_slk, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_stereo, cudaEventDisableTiming) != cudaSuccess) {
      set_error("cudaEventCreate failed");
      r = VO_ERR_CUDA;
    }
  }
  if (r == VO_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) r = VO_ERR_CUDA;
  if (r == VO_OK) r = selfcheck_run(c);
  if (r != VO_OK) {
    vo_destroy(c);
    return r;
  }
  c->opt_host_chains = getenv("VO_B200_SEQ_HOST") != nullptr;      // read per context (tests flip them)
  c->opt_no_pyramid_ahead = getenv("VO_B200_NO_PYRAMID_AHEAD") != nullptr;
  if (getenv("VO_B200_ISLAND") && atoi(getenv("VO_B200_ISLAND")) > 0) {
    r = partition_create(c, atoi(getenv("VO_B200_ISLAND")));
    if (r != VO_OK) {
      vo_destroy(c);
      return r;
    }
    // the stereo chain's LK launch goes to the large partition through its own (low-priority) stream and events
    c->aux->s_lk_aux = c->s_lk_aux;
    for (int i = 0; i < 2; i++)
      if (cudaEventCreateWithFlags(&c->aux->ev_p[i], cudaEventDisableTiming) != cudaSuccess) {
        vo_destroy(c);
        return VO_ERR_CUDA;
      }
  }
  c->opt_lookahead = getenv("VO_B200_LOOKAHEAD") != nullptr;
  c->worker = std::thread(worker_main, c);
  *out = c;
  return VO_OK;
}

int vo_self_check(vo_ctx* c) {
  if (!c) return VO_ERR_INVALID_ARG;
  VO_CUDA(cudaSetDevice(c->device));
  return selfcheck_run(c);
}

int vo_destroy(vo_ctx* c) {
  if (!c) return VO_OK;
  cudaSetDevice(c->device);
  if (c->worker.joinable()) {
    {
      std::lock_guard<std::mutex> g(c->mtx);
      c->quit = true;
      c->cv.notify_all();
    }
    c->worker.join();
  }
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->aux) {
    free_chain(c->aux);
    delete c->aux;
  }
  if (c->la) {
    free_chain(c->la);
    delete c->la;
  }
  partition_destroy(c);
  if (c->ev_la) cudaEventDestroy(c->ev_la);
  if (c->ev_gather) cudaEventDestroy(c->ev_gather);
  if (c->ev_slk) cudaEventDestroy(c->ev_slk);
  if (c->pyr) {
    for (int s = 0; s < 4; s++) pyr_free(c->pyr[s]);
    delete[] c->pyr;
  }
  if (c->ev_left) cudaEventDestroy(c->ev_left);
  if (c->ev_stereo) cudaEventDestroy(c->ev_stereo);
  if (c->ev_lk) cudaEventDestroy(c->ev_lk);
  if (c->ev_xform) cudaEventDestroy(c->ev_xform);
  if (c->copy_stream) {
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamDestroy(c->copy_stream);
  }
  for (int i = 0; i < 2; i++) {
    if (c->ev_prefetch[i]) cudaEventDestroy(c->ev_prefetch[i]);
    cudaFree(c->d_stage[i][0]);
    cudaFree(c->d_stage[i][1]);
  }
  free_chain(c);
  delete c;
  return VO_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------ helpers
#include <chrono>
struct HostTrace {
  bool on = getenv("VO_B200_TRACE_HOST") != nullptr;
  double acc[8] = {0};
  long n[8] = {0};
  ~HostTrace() {
    if (!on) return;
    const char* names[8] = {"pnp_sample_gen", "pnp_enqueue", "pnp_sync_wait", "f_sample_gen", "f_enqueue", "f_sync_wait",
                            "lk_sync_wait", "other"};
    for (int i = 0; i < 8; i++)
      if (n[i]) fprintf(stderr, "[vo trace] %-16s calls %6ld  avg %8.2f us\n", names[i], n[i], acc[i] / n[i]);
  }
};
static HostTrace g_trace;
struct TraceScope {
  int k;
  std::chrono::steady_clock::time_point t0;
  explicit TraceScope(int kk) : k(kk) { if (g_trace.on) t0 = std::chrono::steady_clock::now(); }
  ~TraceScope() {
    if (g_trace.on) {
      g_trace.acc[k] += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
      g_trace.n[k]++;
    }
  }
};

#define CHECK_CTX(c)                        \
  if (!(c)) return VO_ERR_INVALID_ARG;      \
  VO_CUDA(cudaSetDevice((c)->device))

static int sync_stream(vo_ctx* c) {
  VO_CUDA(cudaStreamSynchronize(c->stream));
  return VO_OK;
}

// image -> padded level 0 of `slot` (H2D or D2D straight into place), then the pyramid
static int load_image(vo_ctx* c, int slot, const uint8_t* img, int stride, int is_device, bool with_deriv) {
  if (!img) return VO_ERR_INVALID_ARG;
  const int row_bytes = c->p.width * c->p.channels;
  if (stride < row_bytes) {
    set_error("stride %d < %d bytes per image row", stride, row_bytes);
    return VO_ERR_INVALID_ARG;
  }
  PyrLevel& L0 = c->pyr[slot].lv[0];
  if (c->p.channels == 3) {
    // interleaved BGR: tight staging copy (skipped for tight device images), de-interleave into the planes
    const uint8_t* src = img;
    if (!is_device || stride != row_bytes) {
      VO_CUDA(cudaMemcpy2DAsync(c->d_bgr, row_bytes, img, stride, row_bytes, L0.h,
                                is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
      src = c->d_bgr;
    }
    VO_TRY(pyr_split_bgr(c, slot, src));
    return pyr_build(c, slot, nullptr, with_deriv);
  }
  if (is_device) {
    VO_TRY(pyr_unpack_rows(c, slot, img, stride));
  } else {
    VO_CUDA(cudaMemcpy2DAsync(L0.img + (size_t)PAD_Y * L0.pitch + PAD_L, L0.pitch, img, stride, L0.w, L0.h,
                              cudaMemcpyHostToDevice, c->stream));
  }
  return pyr_build(c, slot, nullptr, with_deriv);
}

static int read_counts(vo_ctx* c) {
  VO_CUDA(cudaMemcpyAsync(c->h_count, c->d_count, 16 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  return sync_stream(c);
}

// --- OpenCV-compatible minimal-sample generation (RANSACPointSetRegistrator::getSubset) ---
static inline void draw_indices(CvRng& rng, int count, int m, int32_t* idx) {
  for (int i = 0; i < m; i++) {
    int v;
    for (;;) {
      v = rng.uniform(0, count);
      bool dup = false;
      for (int k = 0; k < i; k++) dup |= (idx[k] == v);
      if (!dup) break;
    }
    idx[i] = v;
  }
}

static bool have_collinear(const float* pts /*xy interleaved, full set*/, const int32_t* idx, int count) {
  const int i = count - 1;
  const double xi = pts[2 * idx[i]], yi = pts[2 * idx[i] + 1];
  for (int j = 0; j < i; j++) {
    const double dx1 = (double)pts[2 * idx[j]] - xi, dy1 = (double)pts[2 * idx[j] + 1] - yi;
    for (int k = 0; k < j; k++) {
      const double dx2 = (double)pts[2 * idx[k]] - xi, dy2 = (double)pts[2 * idx[k] + 1] - yi;
      if (fabs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2))) return true;
    }
  }
  return false;
}

// Draw accepted 7-subsets into out[k*7..]; returns how many were produced (< want only if
// 10000 attempts failed for one sample, like OpenCV's getSubset).
static int draw_fmat_samples(CvRng& rng, const float* m1, const float* m2, int n, int want, int32_t* out) {
  for (int s = 0; s < want; s++) {
    bool found = false;
    for (int att = 0; att < 10000; att++) {
      int32_t* idx = out + (size_t)s * 7;
      draw_indices(rng, n, 7, idx);
      if (!have_collinear(m1, idx, 7) && !have_collinear(m2, idx, 7)) {
        found = true;
        break;
      }
    }
    if (!found) return s;
  }
  return want;
}

// cv::RANSACUpdateNumIters on the host (the LMedS estimator fixes its sample count with it)
static int host_update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = std::max(p, 0.);
  p = std::min(p, 1.);
  ep = std::max(ep, 0.);
  ep = std::min(ep, 1.);
  double num = std::max(1. - p, DBL_MIN);
  double denom = 1. - pow(1. - ep, (double)model_points);
  if (denom < DBL_MIN) return 0;
  num = log(num);
  denom = log(denom);
  return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)rint(num / denom);
}

// findFundamentalMat(FM_RANSAC) below 15 points (calib3d fundam.cpp): N == 7 -> the raw 7-point result and a mask of
// ones; 8 <= N <= 14 -> OpenCV's LMedS estimator (ransac.cu, fmat_lmeds_kernel).  Same outputs as run_fmat.
static int run_fmat_small(vo_ctx* c, const float2* m1, const float2* m2, int n, double conf, const int32_t* replay,
                          int n_replay, const float* h_m1, const float* h_m2, const std::function<int()>& tail) {
  int H = 1;
  if (n == 7) {
    for (int i = 0; i < 7; i++) c->h_samples[i] = i;
  } else {
    double cf = conf;
    if (cf < DBL_EPSILON || cf > 1 - DBL_EPSILON) cf = 0.99;
    H = replay ? n_replay : host_update_num_iters(cf, 0.45, 7, std::max(c->p.f_max_iters, 1));
    if (H > c->cap_h) {
      set_error("%d LMedS samples exceed max_hypotheses %d", H, c->cap_h);
      return VO_ERR_CAPACITY;
    }
    if (replay) {
      memcpy(c->h_samples, replay, (size_t)H * 7 * sizeof(int32_t));
      for (int i = 0; i < H * 7; i++)
        if (c->h_samples[i] < 0 || c->h_samples[i] >= n) {
          set_error("replay sample index out of range");
          return VO_ERR_INVALID_ARG;
        }
    } else {
      CvRng rng(0xffffffffffffffffULL);
      H = draw_fmat_samples(rng, h_m1, h_m2, n, H, c->h_samples);   // a failed getSubset ends OpenCV's loop
    }
  }
  if (H <= 0) {
    set_error("findFundamentalMat: no admissible 7-point subset");
    return VO_ERR_NO_MODEL;
  }
  VO_CUDA(cudaMemcpyAsync(c->d_samples, c->h_samples, (size_t)H * 7 * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  VO_TRY(fmat_solve_launch(c, m1, m2, c->d_samples, H, c->d_models, c->d_counts));
  if (n == 7) VO_TRY(fmat_seven_launch(c, c->d_counts, n, c->d_sel, c->d_mask));
  else VO_TRY(fmat_lmeds_launch(c, m1, m2, n, c->d_models, c->d_counts, 3 * H, c->d_sel, c->d_mask));
  if (tail) VO_TRY(tail());
  VO_CUDA(cudaMemcpyAsync(c->h_sel, c->d_sel, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  VO_TRY(sync_stream(c));
  c->last_f_h = H;
  return VO_OK;
}

// ------------------------------------------------------------------------------------ F-RANSAC
// Device inputs m1, m2 (n points).  h_m1/h_m2: host copies for the collinearity check (unused
// with a replay list).  On return d_sel holds (best, niters, best_count, records), d_mask the
// best model's inlier mask.
// `tail` (optional) enqueues the consumers of the mask (compaction, count read-back) BEFORE the
// one host synchronisation of a round: OpenCV's adaptive stop almost always ends inside the first
// chunk, so the mask of the first selection is final and no second round trip is needed; when it
// is not, more chunks are evaluated and mask + tail are simply redone.
static int run_fmat(vo_ctx* c, const float2* m1, const float2* m2, int n, double thr, double conf,
                    const int32_t* replay, int n_replay, const float* h_m1, const float* h_m2,
                    const std::function<int()>& tail = nullptr) {
  if (n < 7) {
    // OpenCV returns an empty matrix and leaves the mask untouched; the reference would then index an empty mask
    set_error("findFundamentalMat with %d < 7 points: OpenCV returns no model", n);
    return VO_ERR_TOO_FEW_POINTS;
  }
  if (n < 15) return run_fmat_small(c, m1, m2, n, conf, replay, n_replay, h_m1, h_m2, tail);
  const int max_iters = std::max(c->p.f_max_iters, 1);
  const int H = replay ? n_replay : max_iters;
  if (H > c->cap_h) {
    set_error("%d samples exceed max_hypotheses %d", H, c->cap_h);
    return VO_ERR_CAPACITY;
  }
  const float thr2 = (float)(thr * thr);
  CvRng rng(0xffffffffffffffffULL);
  int done = 0;
  int target = (c->p.f_exhaustive || replay) ? H : std::min(H, 48);
  for (;;) {
    int k = target - done;
    if (k > 0) {
      int32_t* hs = c->h_samples + (size_t)done * 7;
      if (replay) {
        memcpy(hs, replay + (size_t)done * 7, (size_t)k * 7 * sizeof(int32_t));
        for (int i = 0; i < k * 7; i++)
          if (hs[i] < 0 || hs[i] >= n) {
            set_error("replay sample index out of range");
            return VO_ERR_INVALID_ARG;
          }
      } else {
        TraceScope ts(3);
        int got = draw_fmat_samples(rng, h_m1, h_m2, n, k, hs);
        if (got < k) {  // OpenCV: getSubset failed -> the loop ends here
          k = got;
          target = done + k;
        }
      }
      if (k > 0) {
        VO_CUDA(cudaMemcpyAsync(c->d_samples + (size_t)done * 7, hs, (size_t)k * 7 * sizeof(int32_t),
                                cudaMemcpyHostToDevice, c->stream));
        VO_TRY(fmat_solve_launch(c, m1, m2, c->d_samples + (size_t)done * 7, k, c->d_models + (size_t)done * 27,
                                 c->d_counts + (size_t)done * 3));
        VO_TRY(fmat_score_launch(c, m1, m2, n, c->d_models + (size_t)done * 27, c->d_counts + (size_t)done * 3, k,
                                 thr2));
        done += k;
      }
    }
    VO_TRY(select_launch(c, c->d_counts, done, 3, 7, n, conf, max_iters, c->d_sel));
    VO_TRY(fmat_mask_launch(c, m1, m2, n, c->d_models, c->d_sel, thr2, c->d_mask));
    if (tail) VO_TRY(tail());
    VO_CUDA(cudaMemcpyAsync(c->h_sel, c->d_sel, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    {
      TraceScope tw(5);
      VO_TRY(sync_stream(c));
    }
    const int niters = c->h_sel[1];
    if (done >= std::min(niters, H) || k <= 0) break;
    target = std::min(niters, H);
  }
  c->last_f_h = done;
  return VO_OK;
}

// ------------------------------------------------------------------------------------ PnP-RANSAC
static int run_pnp(vo_ctx* c, const float3* xyz, const float2* xy, int n, int iters, double thr, double conf,
                   int min_solver, const int32_t* replay, int n_replay, int* n_inl_out) {
  *n_inl_out = 0;
  if (min_solver != VO_PNP_EPNP5 && min_solver != VO_PNP_P3P4) return VO_ERR_INVALID_ARG;
  if (n < 4) {
    // OpenCV: CV_Assert(npoints >= 4) -- an uncaught cv::Exception in the reference
    set_error("solvePnPRansac needs at least 4 points, got %d", n);
    return VO_ERR_TOO_FEW_POINTS;
  }
  if (min_solver == VO_PNP_P3P4 && n != 4) {
    set_error("VO_PNP_P3P4 is OpenCV's choice for exactly four points (solvePnPRansac with default flags); RANSAC over "
              "P3P samples (flags = SOLVEPNP_P3P, which the reference never passes) is outside this library");
    return VO_ERR_INVALID_ARG;
  }
  if (n <= 5) {
    // the minimal sample is the whole set: OpenCV solves once (P3P for 4, EPnP for 5), keeps every point as an
    // inlier and does not refine (calib3d solvepnp.cpp, `model_points == npoints`)
    VO_TRY(pnp_direct_launch(c, xyz, xy, n, c->d_samples, c->d_models, c->d_counts, c->d_sel, c->d_idx, c->d_count + 4,
                             c->d_pose));
    VO_CUDA(cudaMemcpyAsync(c->h_sel, c->d_sel, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    VO_CUDA(cudaMemcpyAsync(c->h_pose, c->d_pose, 8 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    VO_TRY(read_counts(c));
    c->last_pnp_h = 1;
    if (c->h_sel[0] < 0) {
      set_error("solvePnP: no pose from %d points", n);
      return VO_ERR_NO_MODEL;
    }
    *n_inl_out = c->h_count[4];
    return VO_OK;
  }
  const int max_iters = std::max(iters, 1);
  const int H = replay ? n_replay : max_iters;
  if (H > c->cap_h) {
    set_error("%d samples exceed max_hypotheses %d", H, c->cap_h);
    return VO_ERR_CAPACITY;
  }
  const float thr2 = (float)(thr * thr);
  CvRng rng(0xffffffffffffffffULL);
  int done = 0;
  int target = (c->p.ransac_exhaustive || replay) ? H : std::min(H, 32);
  for (;;) {
    int k = target - done;
    if (k > 0) {
      int32_t* hs = c->h_samples + (size_t)done * 5;
      if (replay) {
        memcpy(hs, replay + (size_t)done * 5, (size_t)k * 5 * sizeof(int32_t));
        for (int i = 0; i < k * 5; i++)
          if (hs[i] < 0 || hs[i] >= n) {
            set_error("replay sample index out of range");
            return VO_ERR_INVALID_ARG;
          }
      } else {
        TraceScope ts(0);
        for (int s = 0; s < k; s++) draw_indices(rng, n, 5, hs + (size_t)s * 5);
      }
      TraceScope te(1);
      VO_CUDA(cudaMemcpyAsync(c->d_samples + (size_t)done * 5, hs, (size_t)k * 5 * sizeof(int32_t),
                              cudaMemcpyHostToDevice, c->stream));
      VO_TRY(pnp_solve_launch(c, xyz, xy, c->d_samples + (size_t)done * 5, k, c->d_models + (size_t)done * 16,
                              c->d_counts + done));
      VO_TRY(pnp_score_launch(c, xyz, xy, n, c->d_models + (size_t)done * 16, c->d_counts + done, k, thr2));
      done += k;
    }
    VO_TRY(select_launch(c, c->d_counts, done, 1, 5, n, conf, max_iters, c->d_sel));
    // consumers of the selection, enqueued before the single synchronisation of this round
    VO_TRY(pnp_mask_launch(c, xyz, xy, n, c->d_models, c->d_sel, thr2, c->d_mask));
    VO_TRY(compact_launch(c, c->d_mask, n, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, c->d_idx, 4));
    VO_TRY(pnp_refine_launch(c, xyz, xy, c->d_idx, c->d_count + 4, c->d_models, c->d_sel, c->d_pose));
    VO_CUDA(cudaMemcpyAsync(c->h_sel, c->d_sel, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    VO_CUDA(cudaMemcpyAsync(c->h_pose, c->d_pose, 8 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    {
      TraceScope tw(2);
      VO_TRY(read_counts(c));
    }
    const int niters = c->h_sel[1];
    if (done >= std::min(niters, H) || k <= 0) break;
    target = std::min(niters, H);
  }
  c->last_pnp_h = done;
  if (c->h_sel[0] < 0) {
    set_error("solvePnPRansac: no hypothesis reached 5 inliers");
    return VO_ERR_NO_MODEL;
  }
  *n_inl_out = c->h_count[4];
  return VO_OK;
}

// Enqueue the host->device copies of the frame announced by vo_seq_prefetch (copy stream, tight staging).
static int prefetch_issue(vo_ctx* c) {
  if (!c->pf_wait_left) return VO_OK;
  const int row_bytes = c->p.width * c->p.channels;
  const int s = c->pf_next;
  VO_CUDA(cudaMemcpy2DAsync(c->d_stage[s][0], row_bytes, c->pf_wait_left, c->pf_wait_stride, row_bytes, c->p.height,
                            cudaMemcpyHostToDevice, c->copy_stream));
  if (c->pf_wait_right)
    VO_CUDA(cudaMemcpy2DAsync(c->d_stage[s][1], row_bytes, c->pf_wait_right, c->pf_wait_stride, row_bytes, c->p.height,
                              cudaMemcpyHostToDevice, c->copy_stream));
  VO_CUDA(cudaEventRecord(c->ev_prefetch[s], c->copy_stream));
  c->pf_left[s] = c->pf_wait_left;
  c->pf_right[s] = c->pf_wait_right;
  c->pf_stride[s] = c->pf_wait_stride;
  c->pf_next = 1 - s;
  c->pf_wait_left = nullptr;
  return VO_OK;
}

// ------------------------------------------------------------------------------------ stage pipelines (device)
// LK slot_a -> slot_b of d_in (n points) + status compaction.  Optional xyz carried along.
// Leaves survivors in d_c_ref / d_c_trk / d_c_xyz; *m = count (host, synchronised).
// pts_to_host: also bring the (at most n) surviving correspondences to c->h_pts in the same
// synchronisation -- the F-matrix sampler needs them for OpenCV's collinearity subset check.
static int lk_and_compact(vo_ctx* c, int slot_a, int slot_b, const float2* d_in, const float3* d_in_xyz, int n, int* m,
                          bool pts_to_host = false, bool lk_done = false) {
  *m = 0;
  if (n <= 0) return VO_OK;
  // lk_done: d_xy_trk / d_status were filled from the look-ahead launch (vo_seq_track)
  if (!lk_done) VO_TRY(lk_launch(c, slot_a, slot_b, d_in, n, c->d_xy_trk, c->d_status, nullptr));   // err is not consumed: test only
  VO_TRY(compact_launch(c, c->d_status, n, d_in, c->d_c_ref, c->d_xy_trk, c->d_c_trk, d_in_xyz, c->d_c_xyz, nullptr, 0));
  if (pts_to_host) {
    VO_CUDA(cudaMemcpyAsync(c->h_pts, c->d_c_ref, (size_t)n * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
    VO_CUDA(cudaMemcpyAsync(c->h_pts + (size_t)2 * c->cap, c->d_c_trk, (size_t)n * sizeof(float2), cudaMemcpyDeviceToHost,
                            c->stream));
  }
  if (!c->is_aux && !c->pf_by_worker) VO_TRY(prefetch_issue(c));   // the next frame's H2D copies go out while this LK runs
  {
    TraceScope tw(6);
    VO_TRY(read_counts(c));
  }
  *m = c->h_count[0];
  return VO_OK;
}

// F-RANSAC on d_c_* (m points) + mask compaction into d_f_*; *k = survivors.
// have_host_pts: c->h_pts already holds the correspondences (lk_and_compact(pts_to_host)).
static int fmat_and_compact(vo_ctx* c, int m, double thr, bool with_xyz, int* k, bool have_host_pts = false,
                            int32_t* idx_out = nullptr) {
  *k = 0;
  if (m <= 0) return VO_OK;
  float* h1 = c->h_pts;
  float* h2 = c->h_pts + (size_t)2 * c->cap;
  if (!have_host_pts) {
    VO_CUDA(cudaMemcpyAsync(h1, c->d_c_ref, (size_t)m * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
    VO_CUDA(cudaMemcpyAsync(h2, c->d_c_trk, (size_t)m * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
    VO_TRY(sync_stream(c));
  }
  auto tail = [&]() -> int {
    VO_TRY(compact_launch(c, c->d_mask, m, c->d_c_ref, c->d_f_ref, c->d_c_trk, c->d_f_trk,
                          with_xyz ? c->d_c_xyz : nullptr, c->d_f_xyz, idx_out, 1));
    VO_CUDA(cudaMemcpyAsync(c->h_count, c->d_count, 16 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    return VO_OK;
  };
  VO_TRY(run_fmat(c, c->d_c_ref, c->d_c_trk, m, thr, c->p.f_conf, nullptr, 0, h1, h2, tail));
  *k = c->h_count[1];
  return VO_OK;
}

static void make_projections(const vo_params& p, double* P /*24*/) {
  // P1 = K [I|0], P2 = K [I | (-b,0,0)^T]   (reference src/triangulation.cpp:142-149)
  const double K[9] = {p.fx, 0, p.cx, 0, p.fy, p.cy, 0, 0, 1};
  double E1[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  double E2[12] = {1, 0, 0, -p.baseline, 0, 1, 0, 0, 0, 0, 1, 0};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 4; j++) {
      double s1 = 0, s2 = 0;
      for (int k = 0; k < 3; k++) {
        s1 += K[i * 3 + k] * E1[k * 4 + j];
        s2 += K[i * 3 + k] * E2[k * 4 + j];
      }
      P[i * 4 + j] = s1;
      P[12 + i * 4 + j] = s2;
    }
}

// stereoTriangulate on pyramids slot_l / slot_r (already built).  Result: d_f_ref (2-D left),
// d_xyz_tmp (camera frame) and, when pose != NULL, d_f_xyz <- pose * xyz (world frame).
// after_lk (optional): called with the number of stereo-LK survivors (d_c_ref) before F-RANSAC is enqueued;
// idx_out (optional): indices of the F-RANSAC inliers in that survivor order.
static int stereo_pipeline(vo_ctx* c, int slot_l, int slot_r, const double* pose3x4, int* n_out, int* n_grid_out,
                           const std::function<int(int)>& after_lk = nullptr, int32_t* idx_out = nullptr) {
  int ng = 0;
  VO_TRY(grid_launch(c, c->p.height, c->p.width, c->p.grid_step, c->d_xy_in, &ng));
  if (n_grid_out) *n_grid_out = ng;
  int m = 0, k = 0;
  VO_TRY(lk_and_compact(c, slot_l, slot_r, c->d_xy_in, nullptr, ng, &m, true));
  if (after_lk) VO_TRY(after_lk(m));
  VO_TRY(fmat_and_compact(c, m, c->p.f_thr_stereo, false, &k, true, idx_out));
  double P[36];
  make_projections(c->p, P);
  if (pose3x4) memcpy(P + 24, pose3x4, 12 * sizeof(double));
  VO_CUDA(cudaMemcpyAsync(c->d_cam, P, (pose3x4 ? 36 : 24) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  VO_TRY(triangulate_launch(c, c->d_cam, c->d_f_ref, c->d_f_trk, k, c->d_xyz_tmp, pose3x4 ? c->d_cam + 24 : nullptr,
                            c->d_f_xyz));
  *n_out = k;
  return VO_OK;
}

// PyrLKtrackFrame2Frame on device data: d_ref_xy/d_ref_xyz (n) -> d_f_trk, d_f_xyz, d_f_ref (k)
static int track_pipeline(vo_ctx* c, int slot_ref, int slot_cur, const float2* d_ref_xy, const float3* d_ref_xyz, int n,
                          int* k, bool lk_done = false) {
  int m = 0;
  VO_TRY(lk_and_compact(c, slot_ref, slot_cur, d_ref_xy, d_ref_xyz, n, &m, true, lk_done));
  return fmat_and_compact(c, m, c->p.f_thr_temporal, true, k, true);
}

// two-attempt PnP on d_f_xyz / d_f_trk (k points); pose in h_pose, inliers in d_idx
static int pnp_two_attempts(vo_ctx* c, int k, int* n_inl, int* attempt) {
  *attempt = 1;
  int r = run_pnp(c, c->d_f_xyz, c->d_f_trk, k, c->p.pnp_iters, c->p.pnp_thr, c->p.pnp_conf, VO_PNP_EPNP5, nullptr, 0,
                  n_inl);
  if (r != VO_OK && r != VO_ERR_NO_MODEL && r != VO_ERR_TOO_FEW_POINTS) return r;
  if (*n_inl < c->p.pnp_min_inliers) {
    *attempt = 2;
    double keep[6];
    const bool had = (r == VO_OK);
    if (had) memcpy(keep, c->h_pose, sizeof(keep));
    r = run_pnp(c, c->d_f_xyz, c->d_f_trk, k, c->p.pnp_retry_iters, c->p.pnp_retry_thr, c->p.pnp_retry_conf,
                VO_PNP_EPNP5, nullptr, 0, n_inl);
    if (r != VO_OK && r != VO_ERR_NO_MODEL && r != VO_ERR_TOO_FEW_POINTS) return r;
    if (r != VO_OK && had) memcpy(c->h_pose, keep, sizeof(keep));
    if (*n_inl < c->p.pnp_min_inliers) return VO_ERR_LOW_INLIERS;
  }
  return VO_OK;
}

// ------------------------------------------------------------------------------------ fused chains
// The same stages as track_pipeline + pnp_two_attempts / stereo_pipeline, but enqueued back to back
// with every element count read from device memory and the RANSAC samples drawn on the device, so
// the chain needs ONE host synchronisation at its end instead of one per stage.  Anything outside
// the common case (adaptive stop not reached inside the first chunk, too few points, low inlier
// count -> second PnP attempt, sampler overflow) is detected after that synchronisation and the
// stage is redone on the host-driven path above, which handles every case.
// First-chunk size of the fused F-RANSAC.  OpenCV stops after niters = log(1-conf)/log(1-w^7) samples
// (w = inlier ratio of the best model so far): 48 samples cover w >= 0.71, which the stereo pairs
// (threshold 3 px) always reach; the temporal pairs (threshold 1 px) sit at w = 0.60..0.70 on ~13 % of
// the bench frames (niters 49..143), and every miss costs a host-driven redo of F + PnP (~1 ms).
// 96 samples cover w >= 0.64 for +14 us of scoring.
static int fused_f_chunk(const vo_ctx* c, bool temporal) {
  if (c->p.f_exhaustive) return std::min(std::max(c->p.f_max_iters, 1), 1024);
  return std::min(temporal ? F_CHUNK_TEMPORAL : F_CHUNK, std::max(c->p.f_max_iters, 1));
}

struct FusedStatus {
  int m = 0, k = 0, n_inl = 0;
  bool f_ok = false, pnp_ok = false;
};

// F-RANSAC part of a fused chain on d_c_* (count in d_count[0], upper bound n_max):
// sampling, solve, score, select (-> d_sel + 4), mask, compaction into d_f_* (count d_count[1]).
static int enqueue_fmat_fused(vo_ctx* c, int n_max, double thr, bool with_xyz) {
  const int H = fused_f_chunk(c, with_xyz);   // with_xyz <=> temporal pair
  const float thr2 = (float)(thr * thr);
  c->n_dev = c->d_count + 0;
  VO_TRY(sample_launch(c, 7, c->d_c_ref, c->d_c_trk, n_max, H, c->d_samples, c->d_flags + 0));
  VO_TRY(fmat_solve_launch(c, c->d_c_ref, c->d_c_trk, c->d_samples, H, c->d_models, c->d_counts));
  VO_TRY(fmat_score_select_launch(c, c->d_c_ref, c->d_c_trk, n_max, c->d_models, c->d_counts, H, thr2, c->p.f_conf,
                                  std::max(c->p.f_max_iters, 1), c->d_sel + 4));
  VO_TRY(fmat_mask_compact_launch(c, c->d_c_ref, c->d_c_trk, with_xyz ? c->d_c_xyz : nullptr, n_max, c->d_models, c->d_sel + 4,
                                  thr2, c->d_mask, c->d_f_ref, c->d_f_trk, c->d_f_xyz, 1));
  c->n_dev = nullptr;
  c->last_f_h = H;
  return VO_OK;
}

static bool fmat_fused_valid(const vo_ctx* c, int m, int H) {
  // h_sel[4..7] = F selection; the first chunk is final iff OpenCV's niters ended inside it
  const bool ok = c->h_flags[0] == 0 && m >= 15 && c->h_sel[4] >= 0 && c->h_sel[5] <= H;
  static const bool dbg = getenv("VO_B200_DEBUG_FALLBACK") != nullptr;
  if (dbg && !ok)
    fprintf(stderr, "[vo fallback] F chain %d: flag %d m %d best %d niters %d count %d (H %d)\n", c->is_aux ? 1 : 0, c->h_flags[0],
            m, c->h_sel[4], c->h_sel[5], c->h_sel[6], H);
  return ok;
}

// d_n (optional): device-resident number of reference points (n is then only the upper bound the
// grids are sized with) -- what a CUDA-graph replay needs.
static int track_pnp_fused_enqueue(vo_ctx* c, int slot_ref, int slot_cur, const float2* d_ref_xy,
                                   const float3* d_ref_xyz, int n, const int* d_n = nullptr, bool lk_done = false) {
  const int iters = std::max(c->p.pnp_iters, 1);
  const int Hp = c->p.ransac_exhaustive ? iters : std::min(iters, PNP_CHUNK);
  const float thr2 = (float)(c->p.pnp_thr * c->p.pnp_thr);
  c->n_dev = d_n;
  if (c->s_island) {
    // SM partition: the LK launch on the large partition, the F-RANSAC stage on the island no LK launch can occupy
    VO_TRY(on_stream(c, c->s_lk, c->ev_p[0], c->ev_p[1], [&]() {
      return lk_launch(c, slot_ref, slot_cur, d_ref_xy, n, c->d_xy_trk, c->d_status, nullptr);
    }));
    // latency-bound kernels (compaction, sampling, 7-point solve, mask + compaction) on the island; the scoring launch
    // is throughput work (1,080 CTAs) and runs on the whole device
    const int H = fused_f_chunk(c, true);
    const float thr2f = (float)(c->p.f_thr_temporal * c->p.f_thr_temporal);
    VO_TRY(on_stream(c, c->s_island, c->ev_p[2], c->ev_p[3], [&]() {
      VO_TRY(compact_launch(c, c->d_status, n, d_ref_xy, c->d_c_ref, c->d_xy_trk, c->d_c_trk, d_ref_xyz, c->d_c_xyz, nullptr, 0));
      c->n_dev = c->d_count + 0;
      VO_TRY(sample_launch(c, 7, c->d_c_ref, c->d_c_trk, n, H, c->d_samples, c->d_flags + 0));
      return fmat_solve_launch(c, c->d_c_ref, c->d_c_trk, c->d_samples, H, c->d_models, c->d_counts);
    }));
    VO_TRY(fmat_score_select_launch(c, c->d_c_ref, c->d_c_trk, n, c->d_models, c->d_counts, H, thr2f, c->p.f_conf,
                                    std::max(c->p.f_max_iters, 1), c->d_sel + 4));
    VO_TRY(on_stream(c, c->s_island, c->ev_p[0], c->ev_p[1], [&]() {
      return fmat_mask_compact_launch(c, c->d_c_ref, c->d_c_trk, c->d_c_xyz, n, c->d_models, c->d_sel + 4, thr2f, c->d_mask,
                                      c->d_f_ref, c->d_f_trk, c->d_f_xyz, 1);
    }));
    c->n_dev = nullptr;
    c->last_f_h = H;
  } else {
  if (!lk_done) VO_TRY(lk_launch(c, slot_ref, slot_cur, d_ref_xy, n, c->d_xy_trk, c->d_status, nullptr));
  if (c->ev_lk_done) VO_CUDA(cudaEventRecord(c->ev_lk_done, c->stream));
  VO_TRY(compact_launch(c, c->d_status, n, d_ref_xy, c->d_c_ref, c->d_xy_trk, c->d_c_trk, d_ref_xyz, c->d_c_xyz, nullptr, 0));
  c->n_dev = nullptr;
  VO_TRY(enqueue_fmat_fused(c, n, c->p.f_thr_temporal, true));
  }
  // PnP on d_f_* (count in d_count[1])
  c->n_dev = c->d_count + 1;
  VO_TRY(sample_launch(c, 5, nullptr, nullptr, n, Hp, c->d_samples, c->d_flags + 1));
  VO_TRY(pnp_solve_launch(c, c->d_f_xyz, c->d_f_trk, c->d_samples, Hp, c->d_models, c->d_counts));
  VO_TRY(pnp_score_select_launch(c, c->d_f_xyz, c->d_f_trk, n, c->d_models, c->d_counts, Hp, thr2, c->p.pnp_conf, iters, c->d_sel));
  VO_TRY(pnp_mask_compact_launch(c, c->d_f_xyz, c->d_f_trk, n, c->d_models, c->d_sel, thr2, c->d_mask, c->d_idx, 4));
  c->n_dev = nullptr;
  VO_TRY(pnp_refine_launch(c, c->d_f_xyz, c->d_f_trk, c->d_idx, c->d_count + 4, c->d_models, c->d_sel, c->d_pose));
  c->last_pnp_h = Hp;
  VO_CUDA(cudaMemcpyAsync(c->h_res, c->d_res, RES_BYTES, cudaMemcpyDeviceToHost, c->stream));   // pose, counts, selection, flags
  return VO_OK;
}

static int track_pnp_fused_finish(vo_ctx* c, FusedStatus* st) {
  if (!c->is_aux) VO_TRY(prefetch_issue(c));
  const int Hf = fused_f_chunk(c, true);
  const int iters = std::max(c->p.pnp_iters, 1);
  const int Hp = c->p.ransac_exhaustive ? iters : std::min(iters, PNP_CHUNK);
  VO_TRY(sync_stream(c));
  st->m = c->h_count[0];
  st->k = c->h_count[1];
  st->n_inl = c->h_count[4];
  st->f_ok = fmat_fused_valid(c, st->m, Hf);
  st->pnp_ok = st->f_ok && c->h_flags[1] == 0 && st->k > 5 && c->h_sel[0] >= 0 && c->h_sel[1] <= Hp &&
               st->n_inl >= c->p.pnp_min_inliers;
  static const bool dbg = getenv("VO_B200_DEBUG_FALLBACK") != nullptr;
  if (dbg && st->f_ok && !st->pnp_ok)
    fprintf(stderr, "[vo fallback] PnP: flag %d k %d best %d niters %d n_inl %d (H %d)\n", c->h_flags[1], st->k, c->h_sel[0],
            c->h_sel[1], st->n_inl, Hp);
  return VO_OK;
}

// stereoTriangulate as one fused chain.  Same outputs as stereo_pipeline (pose == NULL variant).
static int stereo_fused_enqueue(vo_ctx* c, int slot_l, int slot_r, int* n_grid_out, cudaEvent_t lk_after = nullptr) {
  int ng = 0;
  VO_TRY(grid_launch(c, c->p.height, c->p.width, c->p.grid_step, c->d_xy_in, &ng));
  if (n_grid_out) *n_grid_out = ng;
  if (ng <= 0) {
    VO_CUDA(cudaMemsetAsync(c->d_count, 0, 16 * sizeof(int), c->stream));
    VO_CUDA(cudaMemcpyAsync(c->h_count, c->d_count, 16 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    return VO_OK;
  }
  // lk_after: the other chain's LK launch.  Either LK fills every SM; side by side they only slow each
  // other down, back to back this one overlaps the other chain's latency-bound RANSAC solvers instead.
  if (lk_after) VO_CUDA(cudaStreamWaitEvent(c->stream, lk_after, 0));
  if (c->s_lk_aux) {
    VO_TRY(on_stream(c, c->s_lk_aux, c->ev_p[0], c->ev_p[1], [&]() {
      return lk_launch(c, slot_l, slot_r, c->d_xy_in, ng, c->d_xy_trk, c->d_status, nullptr);
    }));
  } else
  VO_TRY(lk_launch(c, slot_l, slot_r, c->d_xy_in, ng, c->d_xy_trk, c->d_status, nullptr));
  VO_TRY(compact_launch(c, c->d_status, ng, c->d_xy_in, c->d_c_ref, c->d_xy_trk, c->d_c_trk, nullptr, nullptr, nullptr, 0));
  VO_TRY(enqueue_fmat_fused(c, ng, c->p.f_thr_stereo, false));
  c->n_dev = c->d_count + 1;
  VO_TRY(triangulate_launch(c, c->d_Pst, c->d_f_ref, c->d_f_trk, ng, c->d_xyz_tmp, nullptr, nullptr));
  c->n_dev = nullptr;
  VO_CUDA(cudaMemcpyAsync(c->h_res, c->d_res, RES_BYTES, cudaMemcpyDeviceToHost, c->stream));   // counts, selection, flags
  return VO_OK;
}

// Synchronise a stereo chain enqueued by stereo_fused_enqueue; when its fused result is not final,
// redo F-RANSAC + triangulation on the host-driven path.
// Result: d_f_ref (left 2-D), d_xyz_tmp (camera frame), *n_out points.
static int stereo_fused_finish(vo_ctx* c, int* n_out) {
  const int Hf = fused_f_chunk(c, false);
  VO_TRY(sync_stream(c));
  *n_out = c->h_count[1];
  if (c->h_count[0] == 0) {
    *n_out = 0;
    return VO_OK;
  }
  if (fmat_fused_valid(c, c->h_count[0], Hf)) return VO_OK;
  int k = 0;
  VO_TRY(fmat_and_compact(c, c->h_count[0], c->p.f_thr_stereo, false, &k));   // d_c_* still hold the LK survivors
  double P[24];
  make_projections(c->p, P);
  VO_CUDA(cudaMemcpyAsync(c->d_cam, P, sizeof(P), cudaMemcpyHostToDevice, c->stream));
  VO_TRY(triangulate_launch(c, c->d_cam, c->d_f_ref, c->d_f_trk, k, c->d_xyz_tmp, nullptr, nullptr));
  *n_out = k;
  return VO_OK;
}

static int stereo_any(vo_ctx* c, int slot_l, int slot_r, int* n_out, int* n_grid_out) {
  VO_TRY(stereo_fused_enqueue(c, slot_l, slot_r, n_grid_out));
  return stereo_fused_finish(c, n_out);
}

static int temporal_finish(vo_ctx* c, int* k, int* n_inl, int* attempt);
static bool temporal_fusable(const vo_ctx* c) {
  const int iters = std::max(c->p.pnp_iters, 1);
  return (c->p.ransac_exhaustive ? iters : std::min(iters, PNP_CHUNK)) <= 1024;
}

// PerspectiveNpointEstimation on device data: fused chain first, host-driven completion otherwise.
// Returns VO_OK or VO_ERR_LOW_INLIERS like pnp_two_attempts; *k tracked points in d_f_*, pose in h_pose.
static int temporal_any(vo_ctx* c, int slot_ref, int slot_cur, const float2* d_ref_xy, const float3* d_ref_xyz, int n,
                        int* k, int* n_inl, int* attempt) {
  *k = 0;
  *n_inl = 0;
  *attempt = 1;
  if (n <= 0) return pnp_two_attempts(c, 0, n_inl, attempt);
  const int iters = std::max(c->p.pnp_iters, 1);
  if ((c->p.ransac_exhaustive ? iters : std::min(iters, PNP_CHUNK)) > 1024) {   // beyond the device sampler
    VO_TRY(track_pipeline(c, slot_ref, slot_cur, d_ref_xy, d_ref_xyz, n, k));
    return pnp_two_attempts(c, *k, n_inl, attempt);
  }
  VO_TRY(track_pnp_fused_enqueue(c, slot_ref, slot_cur, d_ref_xy, d_ref_xyz, n));
  return temporal_finish(c, k, n_inl, attempt);
}

// second half of temporal_any: synchronise the fused chain and complete it on the host path if needed
static int temporal_finish(vo_ctx* c, int* k, int* n_inl, int* attempt) {
  *k = 0;
  *n_inl = 0;
  *attempt = 1;
  FusedStatus st;
  VO_TRY(track_pnp_fused_finish(c, &st));
  if (st.pnp_ok) {
    *k = st.k;
    *n_inl = st.n_inl;
    return VO_OK;
  }
  if (st.f_ok) {
    *k = st.k;
  } else {
    int r = fmat_and_compact(c, st.m, c->p.f_thr_temporal, true, k);   // d_c_* still hold the LK survivors
    if (r != VO_OK && r != VO_ERR_TOO_FEW_POINTS) return r;
  }
  return pnp_two_attempts(c, *k, n_inl, attempt);
}

// left-image pyramid slots rotate 0 -> 1 -> 3 -> 0 (slot 2 is the right image): reference, current, look-ahead
static inline int next_left_slot(int s) { return s == 0 ? 1 : (s == 1 ? 3 : 0); }

// Enqueue the NEXT frame's temporal LK on the look-ahead chain (called by the stereo worker once this frame's
// stereo-LK survivors -- the superset of the new keyframe's points -- are in aux->d_c_ref, m of them).
static int lookahead_enqueue(vo_ctx* c, int m, int slot_cur, int slot_next, const uint8_t* img, int stride, cudaEvent_t img_ready,
                             const uint8_t* identity) {
  vo_ctx* a = c->aux;
  vo_ctx* l = c->la;
  if (m <= 0) return VO_OK;
  VO_CUDA(cudaEventRecord(c->ev_slk, a->stream));               // survivors complete
  VO_CUDA(cudaStreamWaitEvent(l->stream, c->ev_slk, 0));
  VO_CUDA(cudaStreamWaitEvent(l->stream, c->ev_gather, 0));      // the previous look-ahead's tracks have been gathered
  if (img_ready) VO_CUDA(cudaStreamWaitEvent(l->stream, img_ready, 0));
  VO_TRY(load_image(l, slot_next, img, stride, 1, true));
  VO_TRY(lk_launch(l, slot_cur, slot_next, a->d_c_ref, m, l->d_xy_trk, l->d_status, nullptr));
  VO_CUDA(cudaEventRecord(c->ev_la, l->stream));
  c->la_valid = true;
  c->la_left = identity;
  c->la_m = m;
  return VO_OK;
}

static bool lk_order_env() {
  static const bool v = getenv("VO_B200_LK_ORDER") != nullptr;
  return v;
}

// The announced next frame's left pyramid, built ahead on the `la` stream into the slot the next call will use as its
// current image (ev_la marks its completion).  The image is either device-resident (vo_seq_announce) or on its way into
// the staging buffers (vo_seq_prefetch + prefetch_issue).
static int pyramid_ahead_enqueue(vo_ctx* c, int slot_next, const uint8_t* current_id) {
  vo_ctx* l = c->la;
  if (!l || c->opt_lookahead || c->opt_no_pyramid_ahead) return VO_OK;
  const uint8_t* img = nullptr;
  const uint8_t* id = nullptr;
  int stride = 0;
  cudaEvent_t ready = nullptr;
  if (c->ann_left) {
    img = id = c->ann_left;
    stride = c->ann_stride;
    c->ann_left = c->ann_right = nullptr;
  } else if (c->copy_stream) {
    const int sl = 1 - c->pf_next;     // the staging set the last prefetch_issue filled
    if (c->pf_left[sl]) {
      img = c->d_stage[sl][0];
      id = c->pf_left[sl];
      stride = c->p.width * c->p.channels;
      ready = c->ev_prefetch[sl];
    }
  }
  if (!img || id == current_id) return VO_OK;
  if (ready) VO_CUDA(cudaStreamWaitEvent(l->stream, ready, 0));
  VO_TRY(load_image(l, slot_next, img, stride, 1, true));
  VO_CUDA(cudaEventRecord(c->ev_la, l->stream));
  c->pa_valid = true;
  c->pa_left = id;
  c->pa_slot = slot_next;
  return VO_OK;
}

// a call that touches the pyramid slots outside the sequence driver: let an in-flight look-ahead finish first
static void lookahead_quiesce(vo_ctx* c) {
  if (c->la && (c->la_inflight || c->la_valid || c->pa_valid)) cudaStreamSynchronize(c->la->stream);
  c->pa_valid = false;
  c->lk_ahead = false;
  c->la_inflight = c->la_valid = false;
  c->ann_left = c->ann_right = nullptr;
}

static void pose_from_pnp(const double rvec[3], const double tvec[3], double pose[12]) {
  double R[9];
  rodrigues_vec2mat(rvec, R);
  // R <- R^T ; t <- -R^T tvec   (reference src/VisualSLAM.cpp:70-74), pose = [R|t] (:93-97)
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) pose[i * 4 + j] = R[j * 3 + i];
    double s = 0;
    for (int k = 0; k < 3; k++) s += (-R[k * 3 + i]) * tvec[k];
    pose[i * 4 + 3] = s;
  }
}

extern "C" {

// ------------------------------------------------------------------------------------ stage API
int vo_grid_keypoints(vo_ctx* c, int rows, int cols, int step, float* xy, int cap, int* n) {
  CHECK_CTX(c);
  if (!n || step <= 0) return VO_ERR_INVALID_ARG;
  int ng = 0;
  VO_TRY(grid_launch(c, rows, cols, step, c->d_xy_in, &ng));
  *n = ng;
  if (ng > cap) return VO_ERR_CAPACITY;
  if (ng && xy) VO_CUDA(cudaMemcpyAsync(xy, c->d_xy_in, (size_t)ng * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
  return sync_stream(c);
}

int vo_anms(vo_ctx* c, const float* xy, const float* response, int n, int num_keep, int32_t* keep_idx, int cap,
            int* n_keep) {
  CHECK_CTX(c);
  if (!xy || !response || !keep_idx || !n_keep || n < 0) return VO_ERR_INVALID_ARG;
  return anms_launch(c, xy, response, n, num_keep, keep_idx, cap, n_keep);
}

int vo_lk_track(vo_ctx* c, const uint8_t* prev, const uint8_t* next, int stride, const float* prev_xy, int n,
                float* next_xy, uint8_t* status, float* err) {
  CHECK_CTX(c);
  if (!prev || !next || !prev_xy || !next_xy || !status || n < 0) return VO_ERR_INVALID_ARG;
  if (n > c->cap) return VO_ERR_CAPACITY;
  if (n == 0) return VO_OK;
  c->seq_ref_slot = -1;   // the stage entry points reuse the sequence driver's pyramid slots: vo_seq_init again before vo_seq_track
  lookahead_quiesce(c);
  VO_TRY(load_image(c, 0, prev, stride, 0, true));
  VO_TRY(load_image(c, 1, next, stride, 0, false));
  VO_CUDA(cudaMemcpyAsync(c->d_xy_in, prev_xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
  VO_TRY(lk_launch(c, 0, 1, c->d_xy_in, n, c->d_xy_trk, c->d_status, err ? c->d_err : nullptr));
  VO_CUDA(cudaMemcpyAsync(next_xy, c->d_xy_trk, (size_t)n * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
  VO_CUDA(cudaMemcpyAsync(status, c->d_status, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  if (err) VO_CUDA(cudaMemcpyAsync(err, c->d_err, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  return sync_stream(c);
}

// copy a (w+2*pad) x (h+2*pad) window of level L (all planes) to the host in OpenCV's interleaved layout
static int fetch_level(vo_ctx* c, const PyrLevel& L, int cn, int pad, uint8_t* out_level, int16_t* out_deriv) {
  const size_t off = (size_t)(PAD_Y - pad) * L.pitch + (PAD_L - pad);
  const int pw = L.w + 2 * pad, ph = L.h + 2 * pad;
  if (cn == 1) {
    if (out_level)
      VO_CUDA(cudaMemcpy2DAsync(out_level, pw, L.img + off, L.pitch, pw, ph, cudaMemcpyDeviceToHost, c->stream));
    if (out_deriv)
      VO_CUDA(cudaMemcpy2DAsync(out_deriv, (size_t)pw * 4, L.deriv + off, (size_t)L.pitch * 4, (size_t)pw * 4, ph,
                                cudaMemcpyDeviceToHost, c->stream));
    return sync_stream(c);
  }
  std::vector<uint8_t> hi((size_t)pw * ph);
  std::vector<int16_t> hd((size_t)pw * ph * 2);
  for (int k = 0; k < cn; k++) {
    if (out_level) {
      VO_CUDA(cudaMemcpy2DAsync(hi.data(), pw, L.img + k * L.plane + off, L.pitch, pw, ph, cudaMemcpyDeviceToHost, c->stream));
      VO_TRY(sync_stream(c));
      for (size_t i = 0; i < hi.size(); i++) out_level[i * cn + k] = hi[i];
    }
    if (out_deriv) {
      VO_CUDA(cudaMemcpy2DAsync(hd.data(), (size_t)pw * 4, L.deriv + k * L.plane + off, (size_t)L.pitch * 4, (size_t)pw * 4, ph,
                                cudaMemcpyDeviceToHost, c->stream));
      VO_TRY(sync_stream(c));
      // cv::calcSharrDeriv: sample x*cn+k of a row -> (dx, dy) at [2*(x*cn+k)], [2*(x*cn+k)+1]
      for (size_t i = 0; i < (size_t)pw * ph; i++) {
        out_deriv[(i * cn + k) * 2] = hd[2 * i];
        out_deriv[(i * cn + k) * 2 + 1] = hd[2 * i + 1];
      }
    }
  }
  return VO_OK;
}

int vo_debug_pyramid_level(vo_ctx* c, const uint8_t* img, int stride, int level, uint8_t* out_level, int16_t* out_deriv,
                           int* w, int* h) {
  CHECK_CTX(c);
  if (!img) return VO_ERR_INVALID_ARG;
  c->seq_ref_slot = -1;   // the stage entry points reuse the sequence driver's pyramid slots: vo_seq_init again before vo_seq_track
  lookahead_quiesce(c);
  VO_TRY(load_image(c, 0, img, stride, 0, true));
  Pyramid& p = c->pyr[0];
  if (level < 0 || level >= p.nlevels) return VO_ERR_INVALID_ARG;
  PyrLevel& L = p.lv[level];
  if (w) *w = L.w;
  if (h) *h = L.h;
  return fetch_level(c, L, p.cn, 0, out_level, out_deriv);
}

int vo_debug_pyramid_padded(vo_ctx* c, const uint8_t* img, int stride, int level, int pad, uint8_t* out_level,
                            int16_t* out_deriv) {
  CHECK_CTX(c);
  if (!img || pad < 0 || pad > PAD_Y) return VO_ERR_INVALID_ARG;
  c->seq_ref_slot = -1;   // the stage entry points reuse the sequence driver's pyramid slots: vo_seq_init again before vo_seq_track
  lookahead_quiesce(c);
  VO_TRY(load_image(c, 0, img, stride, 0, true));
  Pyramid& p = c->pyr[0];
  if (level < 0 || level >= p.nlevels) return VO_ERR_INVALID_ARG;
  return fetch_level(c, p.lv[level], p.cn, pad, out_level, out_deriv);
}

int vo_fmat_ransac(vo_ctx* c, const float* xy1, const float* xy2, int n, double thr, double conf,
                   const int32_t* samples7, int n_samples, uint8_t* mask, double F[9], int* n_inliers) {
  CHECK_CTX(c);
  if (!xy1 || !xy2 || !mask || n < 0) return VO_ERR_INVALID_ARG;
  if (n > c->cap) return VO_ERR_CAPACITY;
  VO_CUDA(cudaMemcpyAsync(c->d_c_ref, xy1, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
  VO_CUDA(cudaMemcpyAsync(c->d_c_trk, xy2, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
  VO_TRY(run_fmat(c, c->d_c_ref, c->d_c_trk, n, thr, conf, samples7, samples7 ? n_samples : 0, xy1, xy2));
  VO_CUDA(cudaMemcpyAsync(mask, c->d_mask, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  VO_TRY(sync_stream(c));
  const int best = c->h_sel[0];
  if (n_inliers) *n_inliers = best >= 0 ? c->h_sel[2] : 0;
  if (best < 0) {
    set_error("findFundamentalMat: no model with more than 6 inliers");
    return VO_ERR_NO_MODEL;
  }
  if (F) {
    VO_CUDA(cudaMemcpyAsync(F, c->d_models + (size_t)best * 9, 9 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    VO_TRY(sync_stream(c));
  }
  return VO_OK;
}

int vo_triangulate(vo_ctx* c, const double P1[12], const double P2[12], const float* xy1, const float* xy2, int n,
                   float* xyz) {
  CHECK_CTX(c);
  if (!P1 || !P2 || !xy1 || !xy2 || !xyz || n < 0) return VO_ERR_INVALID_ARG;
  if (n > c->cap) return VO_ERR_CAPACITY;
  if (n == 0) return VO_OK;
  double P[24];
  memcpy(P, P1, 96);
  memcpy(P + 12, P2, 96);
  VO_CUDA(cudaMemcpyAsync(c->d_cam, P, sizeof(P), cudaMemcpyHostToDevice, c->stream));
  VO_CUDA(cudaMemcpyAsync(c->d_f_ref, xy1, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
  VO_CUDA(cudaMemcpyAsync(c->d_f_trk, xy2, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
  VO_TRY(triangulate_launch(c, c->d_cam, c->d_f_ref, c->d_f_trk, n, c->d_xyz_tmp, nullptr, nullptr));
  VO_CUDA(cudaMemcpyAsync(xyz, c->d_xyz_tmp, (size_t)n * sizeof(float3), cudaMemcpyDeviceToHost, c->stream));
  return sync_stream(c);
}

int vo_pnp_ransac(vo_ctx* c, const float* xyz, const float* xy, int n, int iters, double thr, double conf,
                  int min_solver, const int32_t* samples, int n_samples, double rvec[3], double tvec[3],
                  int32_t* inliers, int cap, int* n_inl) {
  CHECK_CTX(c);
  if (!xyz || !xy || !rvec || !tvec || n < 0) return VO_ERR_INVALID_ARG;
  if (n > c->cap) return VO_ERR_CAPACITY;
  VO_CUDA(cudaMemcpyAsync(c->d_f_xyz, xyz, (size_t)n * sizeof(float3), cudaMemcpyHostToDevice, c->stream));
  VO_CUDA(cudaMemcpyAsync(c->d_f_trk, xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
  int ni = 0;
  int r = run_pnp(c, c->d_f_xyz, c->d_f_trk, n, iters, thr, conf, min_solver, samples, samples ? n_samples : 0, &ni);
  if (n_inl) *n_inl = ni;
  VO_TRY(r);
  for (int i = 0; i < 3; i++) {
    rvec[i] = c->h_pose[i];
    tvec[i] = c->h_pose[3 + i];
  }
  if (inliers) {
    if (ni > cap) return VO_ERR_CAPACITY;
    VO_CUDA(cudaMemcpyAsync(inliers, c->d_idx, (size_t)ni * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    VO_TRY(sync_stream(c));
  }
  return VO_OK;
}

int vo_debug_last_pnp(vo_ctx* c, double* models, int32_t* counts, int cap_h, int* n_h, int* best, int* n_iters) {
  CHECK_CTX(c);
  const int h = c->last_pnp_h;
  if (n_h) *n_h = h;
  if (best) *best = c->h_sel[0];
  if (n_iters) *n_iters = c->h_sel[1];
  if (h > cap_h && (models || counts)) return VO_ERR_CAPACITY;
  std::vector<double> tmp((size_t)h * 16);
  if (models && h) {
    VO_CUDA(cudaMemcpyAsync(tmp.data(), c->d_models, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    VO_TRY(sync_stream(c));
    for (int i = 0; i < h; i++)
      for (int k = 0; k < 6; k++) models[i * 6 + k] = tmp[(size_t)i * 16 + k];
  }
  if (counts && h) {
    VO_CUDA(cudaMemcpyAsync(counts, c->d_counts, (size_t)h * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    VO_TRY(sync_stream(c));
  }
  return VO_OK;
}

int vo_debug_last_fmat(vo_ctx* c, double* models, int32_t* counts, int cap_h, int* n_h, int* best_sample,
                       int* best_model, int* n_iters) {
  CHECK_CTX(c);
  const int h = c->last_f_h;
  if (n_h) *n_h = h;
  if (best_sample) *best_sample = c->h_sel[0] >= 0 ? c->h_sel[0] / 3 : -1;
  if (best_model) *best_model = c->h_sel[0] >= 0 ? c->h_sel[0] % 3 : -1;
  if (n_iters) *n_iters = c->h_sel[1];
  if (h > cap_h && (models || counts)) return VO_ERR_CAPACITY;
  if (models && h)
    VO_CUDA(cudaMemcpyAsync(models, c->d_models, (size_t)h * 27 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (counts && h)
    VO_CUDA(cudaMemcpyAsync(counts, c->d_counts, (size_t)h * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  return sync_stream(c);
}

int vo_debug_epnp(vo_ctx* c, const float* obj15, const float* img10, double* dbg432) {
  CHECK_CTX(c);
  float* d_in = (float*)c->d_xyz_in;
  double* d_dbg = c->d_models;
  VO_CUDA(cudaMemcpyAsync(d_in, obj15, 15 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  VO_CUDA(cudaMemcpyAsync(d_in + 16, img10, 10 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  VO_TRY(epnp_debug_launch(c, d_in, d_in + 16, d_dbg));
  VO_CUDA(cudaMemcpyAsync(dbg432, d_dbg, 432 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  return sync_stream(c);
}

int vo_transform_points(vo_ctx* c, const double M[12], const float* xyz_in, int n, float* xyz_out) {
  CHECK_CTX(c);
  if (!M || !xyz_in || !xyz_out || n < 0) return VO_ERR_INVALID_ARG;
  if (n > c->cap) return VO_ERR_CAPACITY;
  if (n == 0) return VO_OK;
  VO_CUDA(cudaMemcpyAsync(c->d_cam + 24, M, 96, cudaMemcpyHostToDevice, c->stream));
  VO_CUDA(cudaMemcpyAsync(c->d_xyz_in, xyz_in, (size_t)n * sizeof(float3), cudaMemcpyHostToDevice, c->stream));
  VO_TRY(transform_launch(c, c->d_cam + 24, c->d_xyz_in, n, c->d_xyz_tmp));
  VO_CUDA(cudaMemcpyAsync(xyz_out, c->d_xyz_tmp, (size_t)n * sizeof(float3), cudaMemcpyDeviceToHost, c->stream));
  return sync_stream(c);
}

int vo_bgr_to_gray(vo_ctx* c, const uint8_t* bgr, int stride, int is_device, uint8_t* gray, int gray_stride) {
  CHECK_CTX(c);
  const int w = c->p.width, h = c->p.height;
  if (!bgr || !gray || stride < 3 * w || gray_stride < w) return VO_ERR_INVALID_ARG;
  if (is_device) {
    VO_TRY(bgr2gray_launch(c, bgr, stride, gray, gray_stride));
    return sync_stream(c);
  }
  if (!c->d_bgr) VO_CUDA(cudaMalloc(&c->d_bgr, (size_t)3 * w * h));
  if (!c->d_gray) VO_CUDA(cudaMalloc(&c->d_gray, (size_t)w * h));
  VO_CUDA(cudaMemcpy2DAsync(c->d_bgr, 3 * w, bgr, stride, 3 * w, h, cudaMemcpyHostToDevice, c->stream));
  VO_TRY(bgr2gray_launch(c, c->d_bgr, 3 * w, c->d_gray, w));
  VO_CUDA(cudaMemcpy2DAsync(gray, gray_stride, c->d_gray, w, w, h, cudaMemcpyDeviceToHost, c->stream));
  return sync_stream(c);
}

int vo_sor_cloud(vo_ctx* c, const float* xyz, int n, int mean_k, double stddev_mul, int32_t* keep_idx, int cap,
                 int* n_keep, float* mean_dist) {
  CHECK_CTX(c);
  if (!xyz || !keep_idx || !n_keep || n < 0 || mean_k < 1) return VO_ERR_INVALID_ARG;
  *n_keep = 0;
  // visualSLAM::SORcloud, src/rosFuncs.cpp:11-19: points with -z > 500 never enter the cloud
  std::vector<int32_t> cloud;       // input index of every cloud point
  std::vector<int32_t> finite;      // cloud positions of the finite points (PCL skips the others)
  std::vector<float> pts;
  cloud.reserve(n);
  pts.reserve((size_t)3 * n);
  for (int i = 0; i < n; i++) {
    const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    if (-1 * z > 500) continue;
    if (std::isfinite(x) && std::isfinite(y) && std::isfinite(z)) {
      finite.push_back((int32_t)cloud.size());
      pts.push_back(x);
      pts.push_back(y);
      pts.push_back(z);
    }
    cloud.push_back(i);
  }
  const int m = (int)cloud.size(), mf = (int)finite.size();
  if (mf > c->cap) return VO_ERR_CAPACITY;
  std::vector<float> dist(m, 0.f);
  int valid = 0;
  // pcl::StatisticalOutlierRemoval::applyFilterIndices, first pass: a query whose k+1 neighbours cannot be
  // found keeps distance 0 and is not counted
  if (mf > mean_k) {
    VO_CUDA(cudaMemcpyAsync(c->d_xyz_in, pts.data(), (size_t)mf * sizeof(float3), cudaMemcpyHostToDevice, c->stream));
    VO_TRY(sor_mean_knn_launch(c, c->d_xyz_in, mf, mean_k, c->d_err));
    std::vector<float> dv(mf);
    VO_CUDA(cudaMemcpyAsync(dv.data(), c->d_err, (size_t)mf * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    VO_TRY(sync_stream(c));
    for (int j = 0; j < mf; j++) dist[finite[j]] = dv[j];
    valid = mf;
  }
  // second pass: mean and standard deviation of the mean distances, in double, in cloud order
  double sum = 0, sq_sum = 0;
  for (int j = 0; j < m; j++) {
    const float d = dist[j];
    sum += d;
    sq_sum += d * d;           // float product, as in PCL
  }
  const double mean = sum / (double)valid;
  const double variance = (sq_sum - sum * sum / (double)valid) / ((double)valid - 1);
  const double stddev = sqrt(variance);
  const double thr = mean + stddev_mul * stddev;
  int k = 0;
  for (int j = 0; j < m; j++) {
    if (dist[j] > thr) continue;   // NaN threshold (no valid distance) keeps everything, like PCL
    if (k < cap) keep_idx[k] = cloud[j];
    k++;
  }
  *n_keep = k;
  if (mean_dist) {
    for (int i = 0; i < n; i++) mean_dist[i] = -1.f;   // -1: not part of the cloud
    for (int j = 0; j < m; j++) mean_dist[cloud[j]] = dist[j];
  }
  return k > cap ? VO_ERR_CAPACITY : VO_OK;
}

int vo_pose_from_pnp(const double rvec[3], const double tvec[3], double pose3x4[12]) {
  if (!rvec || !tvec || !pose3x4) return VO_ERR_INVALID_ARG;
  pose_from_pnp(rvec, tvec, pose3x4);
  return VO_OK;
}

// ------------------------------------------------------------------------------------ fused stage entry points
int vo_dense_lk_tracking(vo_ctx* c, const uint8_t* ref_img, const uint8_t* cur_img, int stride, const float* ref_xy,
                         int n, float* ref_out, float* trk_out, int* m) {
  CHECK_CTX(c);
  if (!ref_img || !cur_img || !ref_xy || !ref_out || !trk_out || !m || n < 0) return VO_ERR_INVALID_ARG;
  if (n > c->cap) return VO_ERR_CAPACITY;
  *m = 0;
  if (n == 0) return VO_OK;
  c->seq_ref_slot = -1;   // the stage entry points reuse the sequence driver's pyramid slots: vo_seq_init again before vo_seq_track
  lookahead_quiesce(c);
  VO_TRY(load_image(c, 0, ref_img, stride, 0, true));
  VO_TRY(load_image(c, 1, cur_img, stride, 0, false));
  VO_CUDA(cudaMemcpyAsync(c->d_xy_in, ref_xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
  int k = 0;
  VO_TRY(lk_and_compact(c, 0, 1, c->d_xy_in, nullptr, n, &k));
  *m = k;
  if (k) {
    VO_CUDA(cudaMemcpyAsync(ref_out, c->d_c_ref, (size_t)k * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
    VO_CUDA(cudaMemcpyAsync(trk_out, c->d_c_trk, (size_t)k * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
  }
  return sync_stream(c);
}

int vo_fmat_thresholding(vo_ctx* c, const float* ref_xy, const float* trk_xy, int n, float* ref_out, float* trk_out,
                         int* m) {
  CHECK_CTX(c);
  if (!ref_xy || !trk_xy || !ref_out || !trk_out || !m || n < 0) return VO_ERR_INVALID_ARG;
  if (n > c->cap) return VO_ERR_CAPACITY;
  *m = 0;
  VO_CUDA(cudaMemcpyAsync(c->d_c_ref, ref_xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
  VO_CUDA(cudaMemcpyAsync(c->d_c_trk, trk_xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
  int k = 0;
  VO_TRY(fmat_and_compact(c, n, c->p.f_thr_stereo, false, &k));
  *m = k;
  if (k) {
    VO_CUDA(cudaMemcpyAsync(ref_out, c->d_f_ref, (size_t)k * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
    VO_CUDA(cudaMemcpyAsync(trk_out, c->d_f_trk, (size_t)k * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
  }
  return sync_stream(c);
}

static int stereo_host(vo_ctx* c, const uint8_t* left, const uint8_t* right, int stride, const double* pose,
                       float* xyz_a, float* xy_left, float* xyz_b, int cap, int* n) {
  if (!left || !right || !n) {
    set_error("NULL IMG");  // the reference prints this and returns (src/triangulation.cpp:81-84)
    return VO_ERR_INVALID_ARG;
  }
  *n = 0;
  c->seq_ref_slot = -1;   // the stage entry points reuse the sequence driver's pyramid slots: vo_seq_init again before vo_seq_track
  lookahead_quiesce(c);
  VO_TRY(load_image(c, 0, left, stride, 0, true));
  VO_TRY(load_image(c, 2, right, stride, 0, false));
  int k = 0;
  VO_TRY(stereo_any(c, 0, 2, &k, nullptr));
  if (pose && k) {
    VO_CUDA(cudaMemcpyAsync(c->d_cam + 24, pose, 12 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    VO_TRY(transform_launch(c, c->d_cam + 24, c->d_xyz_tmp, k, c->d_f_xyz));
  }
  *n = k;
  if (k > cap) return VO_ERR_CAPACITY;
  if (k) {
    // xyz_a: primary 3-D output (world if pose given, else camera); xyz_b: camera-frame copy
    if (xyz_a)
      VO_CUDA(cudaMemcpyAsync(xyz_a, pose ? c->d_f_xyz : c->d_xyz_tmp, (size_t)k * sizeof(float3), cudaMemcpyDeviceToHost,
                              c->stream));
    if (xyz_b) VO_CUDA(cudaMemcpyAsync(xyz_b, c->d_xyz_tmp, (size_t)k * sizeof(float3), cudaMemcpyDeviceToHost, c->stream));
    if (xy_left) VO_CUDA(cudaMemcpyAsync(xy_left, c->d_f_ref, (size_t)k * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
  }
  return sync_stream(c);
}

int vo_stereo_triangulate(vo_ctx* c, const uint8_t* left, const uint8_t* right, int stride, float* xyz, float* xy_left,
                          int cap, int* n) {
  CHECK_CTX(c);
  return stereo_host(c, left, right, stride, nullptr, xyz, xy_left, nullptr, cap, n);
}

int vo_insert_keyframe(vo_ctx* c, const uint8_t* left, const uint8_t* right, int stride, const double pose3x4[12],
                       float* xyz_world, float* xy_left, float* xyz_cam, int cap, int* n) {
  CHECK_CTX(c);
  if (!pose3x4) return VO_ERR_INVALID_ARG;
  return stereo_host(c, left, right, stride, pose3x4, xyz_world, xy_left, xyz_cam, cap, n);
}

static int track_host(vo_ctx* c, const uint8_t* ref_img, const uint8_t* cur_img, int stride, const float* ref_xy,
                      const float* ref_xyz, int n, int* k) {
  if (!ref_img || !cur_img || !ref_xy || !ref_xyz || n < 0) return VO_ERR_INVALID_ARG;
  if (n > c->cap) return VO_ERR_CAPACITY;
  *k = 0;
  if (n == 0) return VO_OK;
  c->seq_ref_slot = -1;   // the stage entry points reuse the sequence driver's pyramid slots: vo_seq_init again before vo_seq_track
  lookahead_quiesce(c);
  VO_TRY(load_image(c, 0, ref_img, stride, 0, true));
  VO_TRY(load_image(c, 1, cur_img, stride, 0, false));
  VO_CUDA(cudaMemcpyAsync(c->d_xy_in, ref_xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
  VO_CUDA(cudaMemcpyAsync(c->d_xyz_in, ref_xyz, (size_t)n * sizeof(float3), cudaMemcpyHostToDevice, c->stream));
  return track_pipeline(c, 0, 1, c->d_xy_in, c->d_xyz_in, n, k);
}

static int copy_track_outputs(vo_ctx* c, int k, float* trk_xy, float* trk_xyz, float* ref_xy_inl) {
  if (k) {
    if (trk_xy) VO_CUDA(cudaMemcpyAsync(trk_xy, c->d_f_trk, (size_t)k * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
    if (trk_xyz) VO_CUDA(cudaMemcpyAsync(trk_xyz, c->d_f_xyz, (size_t)k * sizeof(float3), cudaMemcpyDeviceToHost, c->stream));
    if (ref_xy_inl)
      VO_CUDA(cudaMemcpyAsync(ref_xy_inl, c->d_f_ref, (size_t)k * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
  }
  return sync_stream(c);
}

int vo_track_frame(vo_ctx* c, const uint8_t* ref_img, const uint8_t* cur_img, int stride, const float* ref_xy,
                   const float* ref_xyz, int n, float* trk_xy, float* trk_xyz, float* ref_xy_inl, int* n_out) {
  CHECK_CTX(c);
  if (!n_out) return VO_ERR_INVALID_ARG;
  int k = 0;
  VO_TRY(track_host(c, ref_img, cur_img, stride, ref_xy, ref_xyz, n, &k));
  *n_out = k;
  return copy_track_outputs(c, k, trk_xy, trk_xyz, ref_xy_inl);
}

int vo_pnp_frame(vo_ctx* c, const uint8_t* ref_img, const uint8_t* cur_img, int stride, const float* ref_xy,
                 const float* ref_xyz, int n, float* trk_xy, float* trk_xyz, float* ref_xy_inl, int* n_trk,
                 double rvec[3], double tvec[3], int32_t* inliers, int cap_inl, int* n_inl, int* attempt_used) {
  CHECK_CTX(c);
  if (!n_trk || !rvec || !tvec || !n_inl) return VO_ERR_INVALID_ARG;
  int k = 0;
  *n_inl = 0;
  if (!ref_img || !cur_img || !ref_xy || !ref_xyz || n < 0) return VO_ERR_INVALID_ARG;
  if (n > c->cap) return VO_ERR_CAPACITY;
  int ni = 0, att = 1;
  int r;
  if (n == 0) {
    r = pnp_two_attempts(c, 0, &ni, &att);
  } else {
    c->seq_ref_slot = -1;   // the stage entry points reuse the sequence driver's pyramid slots: vo_seq_init again before vo_seq_track
  lookahead_quiesce(c);
    VO_TRY(load_image(c, 0, ref_img, stride, 0, true));
    VO_TRY(load_image(c, 1, cur_img, stride, 0, false));
    VO_CUDA(cudaMemcpyAsync(c->d_xy_in, ref_xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
    VO_CUDA(cudaMemcpyAsync(c->d_xyz_in, ref_xyz, (size_t)n * sizeof(float3), cudaMemcpyHostToDevice, c->stream));
    r = temporal_any(c, 0, 1, c->d_xy_in, c->d_xyz_in, n, &k, &ni, &att);
  }
  *n_trk = k;
  if (r != VO_OK && r != VO_ERR_LOW_INLIERS) return r;
  VO_TRY(copy_track_outputs(c, k, trk_xy, trk_xyz, ref_xy_inl));
  if (attempt_used) *attempt_used = att;
  *n_inl = ni;
  if (r != VO_OK && r != VO_ERR_LOW_INLIERS) return r;
  for (int i = 0; i < 3; i++) {
    rvec[i] = c->h_pose[i];
    tvec[i] = c->h_pose[3 + i];
  }
  if (inliers && ni) {
    if (ni > cap_inl) return VO_ERR_CAPACITY;
    VO_CUDA(cudaMemcpyAsync(inliers, c->d_idx, (size_t)ni * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    VO_TRY(sync_stream(c));
  }
  return r;
}

// ------------------------------------------------------------------------------------ sequence driver
int vo_seq_init(vo_ctx* c, const uint8_t* left, const uint8_t* right, int stride, int is_device, int* n_points) {
  CHECK_CTX(c);
  if (!left || !right) return VO_ERR_INVALID_ARG;
  // a new sequence: forget announced frames (their copies, if any, are left to finish)
  c->pf_left[0] = c->pf_left[1] = c->pf_wait_left = nullptr;
  if (c->copy_stream) VO_CUDA(cudaStreamSynchronize(c->copy_stream));
  lookahead_quiesce(c);
  VO_TRY(load_image(c, 0, left, stride, is_device, true));
  VO_TRY(load_image(c, 2, right, stride, is_device, false));
  int k = 0;
  VO_TRY(stereo_any(c, 0, 2, &k, nullptr));
  if (k) {
    VO_CUDA(cudaMemcpyAsync(c->d_seq_xy, c->d_f_ref, (size_t)k * sizeof(float2), cudaMemcpyDeviceToDevice, c->stream));
    VO_CUDA(cudaMemcpyAsync(c->d_seq_xyz, c->d_xyz_tmp, (size_t)k * sizeof(float3), cudaMemcpyDeviceToDevice, c->stream));
  }
  c->seq_n = k;
  c->seq_ref_slot = 0;
  if (n_points) *n_points = k;
  return sync_stream(c);
}

int vo_seq_prefetch(vo_ctx* c, const uint8_t* left, const uint8_t* right, int stride) {
  CHECK_CTX(c);
  const int row_bytes = c->p.width * c->p.channels;
  if (!left || stride < row_bytes) return VO_ERR_INVALID_ARG;
  if (!c->copy_stream) {
    VO_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      VO_CUDA(cudaEventCreateWithFlags(&c->ev_prefetch[i], cudaEventDisableTiming));
      for (int e = 0; e < 2; e++) VO_CUDA(cudaMalloc(&c->d_stage[i][e], (size_t)row_bytes * c->p.height));
    }
  }
  // Only noted here.  The copies are enqueued by the next vo_seq_track call once its own first kernels are in
  // flight (prefetch_issue), so that their host-side cost hides behind GPU work; a frame announced while no
  // vo_seq_track call follows is copied by vo_seq_track itself when it arrives.
  c->pf_wait_left = left;
  c->pf_wait_right = right;
  c->pf_wait_stride = stride;
  return VO_OK;
}

int vo_seq_announce(vo_ctx* c, const uint8_t* left, const uint8_t* right, int stride, int is_device) {
  if (!is_device) return vo_seq_prefetch(c, left, right, stride);
  CHECK_CTX(c);
  if (!left || stride < c->p.width * c->p.channels) return VO_ERR_INVALID_ARG;
  c->ann_left = left;
  c->ann_right = right;
  c->ann_stride = stride;
  return VO_OK;
}

int vo_seq_track(vo_ctx* c, const uint8_t* left, const uint8_t* right, int stride, int is_device, int force_keyframe,
                 vo_frame_result* out) {
  CHECK_CTX(c);
  if (!left || !out || c->seq_ref_slot < 0) return VO_ERR_INVALID_ARG;
  memset(out, 0, sizeof(*out));
  const uint8_t* left_id = left;     // the pointer the caller announced / passes (before the staging substitution below)
  if (!is_device && c->pf_wait_left == left) c->pf_wait_left = nullptr;   // announced but never issued: copy it now
  if (!is_device) {
    // images announced by vo_seq_prefetch are already (being) copied: use the device staging instead
    for (int s = 0; s < 2; s++) {
      if (c->pf_left[s] == left && c->pf_stride[s] == stride && (!right || c->pf_right[s] == right)) {
        VO_CUDA(cudaStreamWaitEvent(c->stream, c->ev_prefetch[s], 0));
        VO_CUDA(cudaStreamWaitEvent(c->aux->stream, c->ev_prefetch[s], 0));
        left = c->d_stage[s][0];
        if (right) right = c->d_stage[s][1];
        stride = c->p.width * c->p.channels;
        is_device = 1;
        c->pf_left[s] = nullptr;
        c->pf_next = 1 - s;      // this call reads staging set s: the next announced frame goes to the other one
        break;
      }
    }
  }
  const int ref = c->seq_ref_slot, cur = next_left_slot(ref);
  out->n_lk_in = c->seq_n;

  // A keyframe that is known before PnP (caller forces it, or the policy fires on every frame
  // because no inlier count can reach kf_min_inliers) does not depend on this frame's tracking:
  // run the stereo pipeline on the auxiliary chain concurrently with tracking + PnP.
  const bool kf_known = right && (force_keyframe || c->p.kf_min_inliers > c->p.max_points);
  // default: the fused single-synchronisation chains (2 host synchronisations per frame); VO_B200_SEQ_HOST=1 selects the
  // host-driven chains (3-4 synchronisations per chain, host-side sampling)
  // With the reference's keyframe rule (keyframe only when inliers < 200) the single tracking chain runs host-driven:
  // its F-RANSAC often needs more than the fused chain's first chunk of samples (threshold 1 px on a thinning point
  // set), and every such miss costs a redo -- measured 1129 vs 927 frames/s (config 1) and 973 vs 802 (config-2 sizes).
  const bool host_driven = c->opt_host_chains || !kf_known;
  // Look-ahead is opt-in (VO_B200_LOOKAHEAD=1): measured on the bench workload it LOSES (858 vs 881 frames/s) -- a full
  // LK launch occupies every SM's register file, so the latency-bound solver kernels of the tracking chain it was meant
  // to hide under wait for LK blocks to retire (pnp_solve 0.27 -> 0.49 ms).  Kept because it is exact and because it
  // is the right schedule once the chains run on disjoint SM partitions.
  const bool la_enabled = c->opt_lookahead;

  // Look-ahead bookkeeping.  have_la: the previous call already built this image's pyramid (slot `cur`) and tracked
  // the keyframe's points into it on the look-ahead chain.
  bool have_la = false;
  if (c->la_inflight) {
    have_la = kf_known && host_driven && c->la_valid && c->la_left == left_id && c->seq_n > 0;
    VO_CUDA(cudaStreamWaitEvent(c->stream, c->ev_la, 0));   // either way: it writes the slot this call uses
    c->la_inflight = false;
    c->la_valid = false;
  }
  bool have_pa = false;     // the previous call built this image's pyramid ahead (pyramid_ahead_enqueue)
  const bool lk_was_ahead = c->lk_ahead;   // ... and enqueued this frame's tracking LK behind its epilogue
  c->lk_ahead = false;
  if (c->pa_valid) {
    have_pa = !have_la && c->pa_left == left_id && c->pa_slot == cur;
    if (!lk_was_ahead) VO_CUDA(cudaStreamWaitEvent(c->stream, c->ev_la, 0));     // either way: it writes a slot this call may use
    c->pa_valid = false;
  }
  const bool lk_done_ahead = lk_was_ahead && have_pa && c->seq_n == c->lk_ahead_n;
  // with derivatives: this left image is the previous image of this frame's stereo LK and of the
  // next frame's temporal LK (one fused launch builds levels, borders and Scharr planes)
  if (!have_la && !have_pa) VO_TRY(load_image(c, cur, left, stride, is_device, true));
  vo_ctx* a = c->aux;
  int kk = 0, ng = 0;
  int k = 0, ni = 0, att = 1;
  int r;
  bool stereo_synced = false;
  if (have_la) {
    // this frame's tracks: the F-RANSAC inliers of the keyframe (aux->d_idx, order of d_seq_xy) out of the look-ahead
    // launch over all stereo-LK survivors
    VO_TRY(gather_tracks_launch(c, a->d_idx, c->seq_n, c->la->d_xy_trk, c->la->d_status, c->d_xy_trk, c->d_status));
  }
  // (host calls in front of the tracking chain's LK launch are GPU idle time at the frame boundary: only those the launch
  // depends on come first)
  const bool fused_frame = kf_known && c->seq_n > 0 && temporal_fusable(c) && !host_driven;
  if (!fused_frame || la_enabled)
    VO_CUDA(cudaEventRecord(c->ev_gather, c->stream));   // the look-ahead buffers and aux->d_idx are free again
  auto release_aux = [&]() -> int {
    if (c->xform_pending) {   // previous frame's world transform still reads the stereo chain's outputs
      VO_CUDA(cudaStreamWaitEvent(a->stream, c->ev_xform, 0));
      c->xform_pending = false;
    }
    return VO_OK;
  };
  if (!fused_frame) VO_TRY(release_aux());
  // Two drivers for the dual-chain frame.  Measured on the bench workload (B200, 60 frames, round 2):
  //   fused single-sync chains enqueued by this thread (device-side sampling)                  868 frames/s, e2e 876
  //   host-driven chains, stereo chain on the worker thread (2-3 synchronisations per chain)  883 frames/s, e2e 873
  // Equal within the run-to-run spread; the fused form is the default because it needs two host synchronisations
  // per frame instead of seven, which is what keeps eight ranks on one host from disturbing each other.  (Round 1
  // measured 919 vs 956 the other way round: without the host gaps the tracking chain reaches pnp_solve_kernel
  // while the stereo chain's LK is still running.)  The look-ahead lives in the host-driven form only.
  if (fused_frame) {
    // both chains are enqueued by this thread, the critical one (tracking + PnP, high-priority
    // stream) first; one synchronisation per chain at the end
    // cur-left pyramid is complete: the stereo chain waits for this.  With the tracking LK already in the stream (LK-ahead)
    // an event recorded here would fire after that launch; the pyramid's own completion event stands in
    cudaEvent_t left_ready = c->ev_left;
    if (lk_done_ahead) left_ready = c->ev_la;
    else VO_CUDA(cudaEventRecord(c->ev_left, c->stream));
    int rs = VO_OK;
    if (getenv("VO_B200_STEREO_FIRST")) {
      VO_TRY(release_aux());
      VO_CUDA(cudaStreamWaitEvent(a->stream, left_ready, 0));
      rs = load_image(a, 2, right, stride, is_device, false);
      if (rs == VO_OK) rs = stereo_fused_enqueue(a, cur, 2, &ng);
      VO_TRY(track_pnp_fused_enqueue(c, ref, cur, c->d_seq_xy, c->d_seq_xyz, c->seq_n));
    } else {
      static const bool lk_order = lk_order_env();   // measured: slower (835 vs 853 frames/s)
      c->ev_lk_done = lk_order ? c->ev_lk : nullptr;
      const int rt = track_pnp_fused_enqueue(c, ref, cur, c->d_seq_xy, c->d_seq_xyz, c->seq_n, nullptr, lk_done_ahead);
      c->ev_lk_done = nullptr;
      VO_TRY(rt);
      VO_TRY(release_aux());
      VO_CUDA(cudaStreamWaitEvent(a->stream, left_ready, 0));
      rs = load_image(a, 2, right, stride, is_device, false);
      if (rs == VO_OK) rs = stereo_fused_enqueue(a, cur, 2, &ng, lk_order ? c->ev_lk : nullptr);
    }
    // everything of this frame is enqueued: the next frame's copies (if announced from the host) and its left pyramid
    VO_TRY(prefetch_issue(c));
    VO_TRY(pyramid_ahead_enqueue(c, next_left_slot(cur), left_id));
    r = temporal_finish(c, &k, &ni, &att);
    if (rs == VO_OK) rs = stereo_fused_finish(a, &kk);
    else cudaStreamSynchronize(a->stream);
    stereo_synced = true;     // the host has waited for the stereo chain: the epilogue needs no event on it
    if (r == VO_OK) r = rs;
  } else {
    if (kf_known) {
      VO_CUDA(cudaEventRecord(c->ev_left, c->stream));
      VO_CUDA(cudaStreamWaitEvent(a->stream, c->ev_left, 0));
      VO_CUDA(cudaStreamWaitEvent(a->stream, c->ev_gather, 0));   // aux->d_idx / d_c_ref of the previous keyframe are released
      // The next frame, if the caller announced it: device-resident (vo_seq_announce) or host (vo_seq_prefetch, copied
      // to the staging buffers by the worker).  Its temporal LK is enqueued by the worker as soon as the stereo-LK
      // survivors are known (lookahead_enqueue).
      const bool do_la = la_enabled && host_driven && c->la != nullptr;
      const uint8_t* nxt_dev = do_la ? c->ann_left : nullptr;
      const int nxt_dev_stride = c->ann_stride;
      const bool nxt_host = do_la && !nxt_dev && c->pf_wait_left != nullptr && c->pf_wait_left != left_id;
      c->ann_left = c->ann_right = nullptr;
      c->pf_by_worker = nxt_host;
      const int la_slot = next_left_slot(cur);
      post_task(c, [=, &kk, &ng]() -> int {
        const uint8_t* nimg = nxt_dev;
        int nstride = nxt_dev_stride;
        cudaEvent_t nready = nullptr;
        const uint8_t* nid = nxt_dev;
        if (nxt_host) {
          VO_TRY(prefetch_issue(c));                 // H2D copies of the announced frame (copy stream)
          const int sl = 1 - c->pf_next;             // the staging set they went to
          nimg = c->d_stage[sl][0];
          nstride = c->p.width * c->p.channels;
          nready = c->ev_prefetch[sl];
          nid = c->pf_left[sl];
        }
        VO_TRY(load_image(a, 2, right, stride, is_device, false));
        if (host_driven) {
          std::function<int(int)> hook = nullptr;
          if (nimg) hook = [=](int m) -> int { return lookahead_enqueue(c, m, cur, la_slot, nimg, nstride, nready, nid); };
          VO_TRY(stereo_pipeline(a, cur, 2, nullptr, &kk, &ng, hook, a->d_idx));
        } else {
          VO_TRY(stereo_any(a, cur, 2, &kk, &ng));             // camera-frame xyz in a->d_xyz_tmp
        }
        VO_CUDA(cudaEventRecord(c->ev_stereo, a->stream));
        return VO_OK;
      });
    }
    if (host_driven) {
      r = track_pipeline(c, ref, cur, c->d_seq_xy, c->d_seq_xyz, c->seq_n, &k, have_la);
      if (r == VO_OK) r = pnp_two_attempts(c, k, &ni, &att);
    } else {
      r = temporal_any(c, ref, cur, c->d_seq_xy, c->d_seq_xyz, c->seq_n, &k, &ni, &att);
    }
    if (kf_known) {
      const int rs = wait_task(c);     // always join: the auxiliary chain must be idle on return
      c->pf_by_worker = false;
      c->la_inflight = c->la_valid;    // set by lookahead_enqueue on the worker thread (ordered by the join)
      if (r == VO_OK) r = rs;
    }
  }
  out->n_tracked = k;
  out->n_inliers = ni;
  out->attempt_used = att;
  if (r != VO_OK) return r;  // incl. VO_ERR_LOW_INLIERS: the reference breaks out of its loop here
  for (int i = 0; i < 3; i++) {
    out->rvec[i] = c->h_pose[i];
    out->tvec[i] = c->h_pose[3 + i];
  }
  pose_from_pnp(out->rvec, out->tvec, out->pose3x4);
  if (kf_known) {
    // insertKeyFrames epilogue: world points = pose * camera points (src/keyFrameManagement.cpp:20-30)
    if (!stereo_synced) VO_CUDA(cudaStreamWaitEvent(c->stream, c->ev_stereo, 0));
    VO_TRY(keyframe_epilogue_launch(c, out->pose3x4, a->d_xyz_tmp, a->d_f_ref, kk, c->d_seq_xyz, c->d_seq_xy));
    c->seq_n = kk;
    out->keyframe = 1;
    out->n_kf_points = kk;
    out->n_lk_in_stereo = ng;
    // Everything the caller receives is already on the host; the world transform of the new keyframe
    // points only feeds the NEXT frame's PnP (stream order), so the call does not wait for it.  The
    // stereo chain must not overwrite its outputs before this has read them: ev_xform.
    VO_CUDA(cudaEventRecord(c->ev_xform, c->stream));
    c->xform_pending = true;
    c->seq_ref_slot = cur;
    // LK-ahead: the announced next frame's pyramid is (being) built; its tracking LK only needs that pyramid and the
    // points the epilogue above has just produced.  Enqueued now, it runs while the caller turns around.
    if (fused_frame && c->pa_valid && c->pa_slot == next_left_slot(cur) && kk > 0 && !c->s_island && !lk_order_env()) {
      VO_CUDA(cudaStreamWaitEvent(c->stream, c->ev_la, 0));
      VO_TRY(lk_launch(c, cur, c->pa_slot, c->d_seq_xy, kk, c->d_xy_trk, c->d_status, nullptr));
      c->lk_ahead = true;
      c->lk_ahead_n = kk;
    }
    return VO_OK;
  } else if (ni < c->p.kf_min_inliers || force_keyframe) {
    // keyframe: src/VisualSLAM.cpp:120-137 -> insertKeyFrames (src/keyFrameManagement.cpp:9-31)
    if (!right) {
      set_error("keyframe required (inliers %d < %d) but no right image was supplied", ni, c->p.kf_min_inliers);
      return VO_ERR_INVALID_ARG;
    }
    VO_TRY(load_image(c, 2, right, stride, is_device, false));
    VO_TRY(stereo_any(c, cur, 2, &kk, &ng));
    VO_CUDA(cudaMemcpyAsync(c->d_cam + 24, out->pose3x4, 12 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (kk) {
      VO_TRY(transform_launch(c, c->d_cam + 24, c->d_xyz_tmp, kk, c->d_seq_xyz));
      VO_CUDA(cudaMemcpyAsync(c->d_seq_xy, c->d_f_ref, (size_t)kk * sizeof(float2), cudaMemcpyDeviceToDevice, c->stream));
    }
    c->seq_n = kk;
    out->keyframe = 1;
    out->n_kf_points = kk;
    out->n_lk_in_stereo = ng;
  } else {
    // ref3dCoords = trked3dCoords; ref2dFeatures = trked2dPts   (src/VisualSLAM.cpp:143-146)
    if (k) {
      VO_CUDA(cudaMemcpyAsync(c->d_seq_xy, c->d_f_trk, (size_t)k * sizeof(float2), cudaMemcpyDeviceToDevice, c->stream));
      VO_CUDA(cudaMemcpyAsync(c->d_seq_xyz, c->d_f_xyz, (size_t)k * sizeof(float3), cudaMemcpyDeviceToDevice, c->stream));
    }
    c->seq_n = k;
  }
  c->seq_ref_slot = cur;  // referenceImg = currentImage (src/VisualSLAM.cpp:151)
  return sync_stream(c);
}

int vo_seq_get_reference(vo_ctx* c, float* xy, float* xyz, int cap, int* n) {
  CHECK_CTX(c);
  if (n) *n = c->seq_n;
  if (c->seq_n > cap && (xy || xyz)) return VO_ERR_CAPACITY;
  if (c->seq_n) {
    if (xy) VO_CUDA(cudaMemcpyAsync(xy, c->d_seq_xy, (size_t)c->seq_n * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
    if (xyz) VO_CUDA(cudaMemcpyAsync(xyz, c->d_seq_xyz, (size_t)c->seq_n * sizeof(float3), cudaMemcpyDeviceToHost, c->stream));
  }
  return sync_stream(c);
}

// ------------------------------------------------------------------------------------ harness helpers
void* vo_cuda_stream(vo_ctx* c) { return c ? (void*)c->stream : nullptr; }

int vo_sync(vo_ctx* c) {
  CHECK_CTX(c);
  VO_TRY(sync_stream(c->la));
  return sync_stream(c);
}

int vo_profile_enable(vo_ctx* c, int mask) {
  CHECK_CTX(c);
  VO_TRY(sync_stream(c));
  VO_TRY(sync_stream(c->aux));
  VO_TRY(sync_stream(c->la));
  prof_drain(c);
  prof_drain(c->aux);
  prof_drain(c->la);
  c->prof.mask = (unsigned)mask;
  c->aux->prof.mask = (unsigned)mask;
  c->la->prof.mask = (unsigned)mask;
  return VO_OK;
}

int vo_profile_read(vo_ctx* c, int kernel, int64_t* launches, double* ms, int reset) {
  CHECK_CTX(c);
  if (kernel < 0 || kernel >= VO_K_COUNT) return VO_ERR_INVALID_ARG;
  VO_TRY(sync_stream(c));
  VO_TRY(sync_stream(c->aux));
  VO_TRY(sync_stream(c->la));
  prof_drain(c);
  prof_drain(c->aux);
  prof_drain(c->la);
  if (launches) *launches = c->prof.launches[kernel] + c->aux->prof.launches[kernel] + c->la->prof.launches[kernel];
  if (ms) *ms = c->prof.ms[kernel] + c->aux->prof.ms[kernel] + c->la->prof.ms[kernel];
  if (reset) {
    c->prof.launches[kernel] = c->aux->prof.launches[kernel] = c->la->prof.launches[kernel] = 0;
    c->prof.ms[kernel] = c->aux->prof.ms[kernel] = c->la->prof.ms[kernel] = 0;
  }
  return VO_OK;
}

int vo_debug_timeline(vo_ctx* c, float* rows, int cap, int* n) {
  CHECK_CTX(c);
  if (!rows || !n) return VO_ERR_INVALID_ARG;
  VO_TRY(sync_stream(c));
  VO_TRY(sync_stream(c->aux));
  VO_TRY(sync_stream(c->la));
  *n = 0;
  cudaEvent_t t0 = nullptr;
  if (!c->prof.pending.empty()) t0 = c->prof.pending.front().a;
  if (!t0) return VO_OK;
  int chain = 0;
  for (vo_ctx* k : {c, c->aux, c->la}) {
    for (auto& pe : k->prof.pending) {
      if (*n >= cap) break;
      float st = 0, du = 0;
      cudaEventSynchronize(pe.b);
      cudaEventElapsedTime(&st, t0, pe.a);
      cudaEventElapsedTime(&du, pe.a, pe.b);
      float* r = rows + 4 * (*n);
      r[0] = (float)chain; r[1] = (float)pe.kind; r[2] = st; r[3] = du;
      (*n)++;
    }
    chain++;
  }
  return VO_OK;
}

int64_t vo_launch_count(vo_ctx* c) {
  return c ? c->launch_count + (c->aux ? c->aux->launch_count : 0) + (c->la ? c->la->launch_count : 0) : 0;
}

int vo_lk_work(vo_ctx* c, int64_t* point_levels, int64_t* iterations) {
  CHECK_CTX(c);
  int64_t pl = 0, it = 0;
  for (vo_ctx* k : {c, c->aux, c->la}) {
    VO_CUDA(cudaMemcpyAsync(k->h_lk_work, k->d_lk_work, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, k->stream));
    VO_TRY(sync_stream(k));
    pl += (int64_t)k->h_lk_work[0];
    it += (int64_t)k->h_lk_work[1];
  }
  if (point_levels) *point_levels = pl;
  if (iterations) *iterations = it;
  return VO_OK;
}

int vo_measure_int32_peak(vo_ctx* c, double* tops) {
  CHECK_CTX(c);
  if (!tops) return VO_ERR_INVALID_ARG;
  return int32_peak_launch(c, tops);
}

int vo_lk_slow_paths(vo_ctx* c, int64_t* window_sums, int64_t* iterations) {
  CHECK_CTX(c);
  int64_t a = 0, b = 0;
  for (vo_ctx* k : {c, c->aux, c->la}) {
    VO_CUDA(cudaMemcpyAsync(k->h_lk_work, k->d_lk_work, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, k->stream));
    VO_TRY(sync_stream(k));
    a += (int64_t)k->h_lk_work[2];
    b += (int64_t)k->h_lk_work[3];
  }
  if (window_sums) *window_sums = a;
  if (iterations) *iterations = b;
  return VO_OK;
}

int vo_measure_fp32_peak(vo_ctx* c, double* tflops) {
  CHECK_CTX(c);
  if (!tflops) return VO_ERR_INVALID_ARG;
  return fp32_peak_launch(c, tflops);
}

int vo_synth_render_dev(vo_ctx* c, int seed, int frame, int eye, uint8_t* out_dev) {
  CHECK_CTX(c);
  if (!out_dev) return VO_ERR_INVALID_ARG;
  return synth_launch(c, seed, frame, eye, out_dev);
}

int vo_alloc_host(void** p, uint64_t bytes) {
  if (!p) return VO_ERR_INVALID_ARG;
  VO_CUDA(cudaMallocHost(p, bytes));
  return VO_OK;
}
int vo_free_host(void* p) {
  VO_CUDA(cudaFreeHost(p));
  return VO_OK;
}
int vo_alloc_dev(vo_ctx* c, void** p, uint64_t bytes) {
  CHECK_CTX(c);
  if (!p) return VO_ERR_INVALID_ARG;
  VO_CUDA(cudaMalloc(p, bytes));
  return VO_OK;
}
int vo_free_dev(vo_ctx* c, void* p) {
  CHECK_CTX(c);
  VO_CUDA(cudaFree(p));
  return VO_OK;
}
int vo_memcpy_d2h(vo_ctx* c, void* dst, const void* src, uint64_t bytes) {
  CHECK_CTX(c);
  VO_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  return sync_stream(c);
}
int vo_memcpy_h2d(vo_ctx* c, void* dst, const void* src, uint64_t bytes) {
  CHECK_CTX(c);
  VO_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  return sync_stream(c);
}

}  // extern "C"
