// common.cuh -- context, device buffers and launch helpers shared by the kernels of
// libvo_b200.so.  Host side of the drop-in boundary declared in include/vo_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/vo_b200.h"

namespace vo {

// ------------------------------------------------------------------------------------
// HBM layout of one pyramid level.  Images are stored PADDED so that the LK window and
// the pyrDown taps never need bounds logic: border PAD_Y rows above/below and PAD_L /
// >= PAD_R columns left/right, filled with BORDER_REFLECT_101 (image) or zeros
// (derivative), exactly the borders cv::buildOpticalFlowPyramid creates with
// winSize = 21.  PAD_L = 32 keeps interior rows 32-byte aligned (pitch % 128 == 0).
constexpr int LK_WIN = 21;
constexpr int PAD_Y = 21;
constexpr int PAD_L = 32;
constexpr int PAD_R = 32;
constexpr int MAX_LEVELS = 4;

// 3-channel (BGR) images are stored PLANAR: `cn` planes per level, each with the 1-channel geometry,
// plane k at img + k*plane / deriv + k*plane.  pyrDown, Scharr and the bilinear window samples are
// per-channel operations and the LK sums run over all channels of the window, so planes give the
// same numbers as OpenCV's interleaved layout while every kernel keeps its 1-channel access pattern.
constexpr int MAX_CN = 3;

struct PyrLevel {
  int w, h;          // interior size
  int pitch;         // bytes per padded image row (multiple of 128)
  size_t plane;      // elements (bytes of img, short2 of deriv) per plane = (h + 2*PAD_Y) * pitch
  uint8_t* img;      // padded buffer base; pixel (x,y) of plane k at img[k*plane + (y+PAD_Y)*pitch + x + PAD_L]
  short2* deriv;     // padded (dx,dy) buffer, same geometry, element pitch = pitch
  uint8_t* img_alloc;  // cudaMalloc'ed block: GUARD_ROWS rows before img and after the last plane, so that the LK
                       // tiles (staged with a margin, lk.cu) may read a few rows beyond the padded image
};
constexpr int GUARD_ROWS = 8;

struct Pyramid {
  PyrLevel lv[MAX_LEVELS];
  int nlevels = 0;
  int cn = 1;
  bool has_deriv = false;
  uint64_t stamp = 0;  // content tag (0 = empty)
  int* d_mono = nullptr;   // device flag (3-channel images): 1 = the three planes are identical (a gray image read as BGR)
  // CUtensorMap (TMA descriptor) of the level-0 interior of each plane: u8, dims (w, h), row stride = pitch
  alignas(64) unsigned char tmap0[MAX_CN][128] = {{0}};
};

struct PyrLevelView {  // what kernels see
  const uint8_t* img;
  const short2* deriv;
  int w, h, pitch;
  unsigned plane;    // elements per plane
};

struct PyrView {
  PyrLevelView lv[MAX_LEVELS];
  int nlevels;
  const int* mono;   // see Pyramid::d_mono
};

struct Profile {
  unsigned mask = 0;  // bit k: bracket launches of kind k with events
  int64_t launches[VO_K_COUNT] = {0};
  double ms[VO_K_COUNT] = {0};
  struct Pending { int kind; cudaEvent_t a, b; };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> pool;
};

}  // namespace vo

struct vo_ctx {
  vo_params p;
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  int64_t launch_count = 0;
  vo::Profile prof;

  // A context is one CHAIN of scratch buffers + one stream.  The public context owns a second,
  // auxiliary chain (`aux`) that shares the pyramids: when a keyframe is known in advance
  // (force_keyframe, or a keyframe policy that fires on every frame) the stereo pipeline runs on
  // the auxiliary chain (own stream, own host worker thread) concurrently with temporal
  // tracking + PnP, whose long latency-bound solver kernels leave the SMs mostly idle.
  vo_ctx* aux = nullptr;
  bool is_aux = false;
  std::thread worker;
  std::mutex mtx;
  std::condition_variable cv;
  std::function<int()> task;
  bool task_pending = false, task_done = false, quit = false;
  int task_result = 0;
  std::string task_error;
  cudaEvent_t ev_left = nullptr, ev_stereo = nullptr, ev_lk = nullptr, ev_xform = nullptr;
  bool xform_pending = false;          // the last keyframe transform was enqueued without a final synchronisation
  cudaEvent_t ev_lk_done = nullptr;    // when set, track_pnp_fused_enqueue records it right after its LK launch

  // Look-ahead (vo_seq_announce / vo_seq_prefetch + a keyframe on every frame): the NEXT frame's temporal LK does
  // not depend on this frame's pose, only on this keyframe's 2-D points -- it runs on a third chain (`la`) as soon
  // as the stereo LK of this frame has its survivors, under the RANSAC solvers of the tracking chain, on ALL
  // survivors; the next call gathers the F-RANSAC inliers' tracks from it (LK is independent per point, so the
  // numbers are those of the in-frame launch).
  vo_ctx* la = nullptr;
  bool la_inflight = false;          // a look-ahead was enqueued during the previous vo_seq_track call
  bool la_valid = false;             // ... and la->d_xy_trk / d_status hold the tracks of `la_left`
  const uint8_t* la_left = nullptr;  // identity of that image: the pointer the caller announced
  int la_m = 0;                      // points it ran on (stereo-LK survivors, order of aux->d_c_ref)
  cudaEvent_t ev_la = nullptr, ev_gather = nullptr, ev_slk = nullptr;
  const uint8_t* ann_left = nullptr;  // frame announced as DEVICE-resident (vo_seq_announce)
  const uint8_t* ann_right = nullptr;
  int ann_stride = 0;
  // Pyramid-ahead (fused two-chain frame): the announced next frame's LEFT pyramid depends on nothing but the image; it
  // is built on the `la` stream while this frame's chains run, and the next call finds it in its `cur` slot
  bool pa_valid = false;
  const uint8_t* pa_left = nullptr;  // identity (the pointer the caller announced / will pass)
  int pa_slot = -1;
  bool opt_no_pyramid_ahead = false; // VO_B200_NO_PYRAMID_AHEAD at vo_create
  // ... and when that pyramid exists, the NEXT frame's tracking LK launch is enqueued at the end of this call, behind the
  // keyframe epilogue that produces its input points: it runs while the caller turns around instead of after it
  bool lk_ahead = false;
  int lk_ahead_n = 0;
  // SM partition (VO_B200_ISLAND=<SMs> at vo_create; CUDA green contexts): an island of a few SMs that no LK launch can
  // occupy runs the tracking chain's F-RANSAC stage, the LK launches of both chains run on the rest
  void* g_island = nullptr;          // CUgreenCtx
  void* g_big = nullptr;
  cudaStream_t s_island = nullptr, s_lk = nullptr, s_lk_aux = nullptr;
  cudaEvent_t ev_p[4] = {nullptr, nullptr, nullptr, nullptr};
  int island_sms = 0;
  bool pf_by_worker = false;         // this call's prefetch copies are issued by the stereo worker thread
  bool opt_host_chains = false;      // VO_B200_SEQ_HOST at vo_create: host-driven chains also when the keyframe is known
  bool opt_lookahead = false;        // VO_B200_LOOKAHEAD at vo_create (needs the host-driven chains)

  // image staging + pyramids: slots 0/1/3 rotate through the left images (reference, current, look-ahead), slot 2 right
  uint8_t* d_raw[3] = {nullptr, nullptr, nullptr};
  vo::Pyramid* pyr = nullptr;   // 4 slots, owned by the primary chain, shared with the other chains
  uint64_t stamp_counter = 0;

  // point buffers (capacity max_points)
  int cap = 0;
  float2 *d_xy_in = nullptr, *d_xy_trk = nullptr;        // LK input / output
  uint8_t* d_status = nullptr;
  float* d_err = nullptr;
  float3* d_xyz_in = nullptr;
  float2 *d_c_ref = nullptr, *d_c_trk = nullptr;         // after status compaction
  float3* d_c_xyz = nullptr;
  float2 *d_f_ref = nullptr, *d_f_trk = nullptr;         // after F-mask compaction
  float3* d_f_xyz = nullptr;
  float3* d_xyz_tmp = nullptr;                           // triangulation / transform output
  uint8_t* d_mask = nullptr;
  int32_t* d_idx = nullptr;                              // inlier indices
  uint8_t* d_res = nullptr;                              // result block: pose (16 doubles) | d_count (16 ints) | d_sel (8) | d_flags (8)
  uint8_t* h_res = nullptr;                              // pinned mirror of it (h_pose, h_count, h_sel, h_flags point into it)
  int* d_count = nullptr;                                // small int scratch (16 ints)
  int* h_count = nullptr;                                // pinned mirror
  unsigned long long* d_tile_state = nullptr;            // compaction look-back (epoch<<32 | tile total)
  unsigned* d_epoch = nullptr;                           // [0] device-side launch epoch of the compaction, [1] CTA completion counter of a score launch
  double* d_Pst = nullptr;                               // P1 | P2 of the stereo rig (constant)
  // When set, kernels take their element count from this device pointer (clamped to the host-side
  // upper bound they were launched with): lets a whole chain run without a host round trip.
  const int* n_dev = nullptr;
  uint32_t* d_rng = nullptr;     // OpenCV's RNG output stream for seed 2^64-1 (fixed), RNG_LEN values
  int* d_flags = nullptr;        // sampler status words (0 = ok)
  int* h_flags = nullptr;

  // vo_seq_prefetch: the next frame's images are copied to tight device staging (two sets, alternating) on a
  // copy stream while the current frame is being processed
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_prefetch[2] = {nullptr, nullptr};
  uint8_t* d_stage[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [set][left/right]
  const uint8_t* pf_left[2] = {nullptr, nullptr};
  const uint8_t* pf_right[2] = {nullptr, nullptr};
  int pf_stride[2] = {0, 0};
  int pf_next = 0;
  const uint8_t* pf_wait_left = nullptr;   // announced, copies not enqueued yet
  const uint8_t* pf_wait_right = nullptr;
  int pf_wait_stride = 0;

  // sequence state (vo_seq_*): reference set resident in HBM
  float2* d_seq_xy = nullptr;
  float3* d_seq_xyz = nullptr;
  int seq_n = 0;
  int seq_ref_slot = -1;

  // RANSAC buffers (capacity max_hypotheses)
  int cap_h = 0;
  int32_t* d_samples = nullptr;   // H x 7
  int32_t* h_samples = nullptr;   // pinned
  double* d_models = nullptr;     // F: H x 3 x 9 ; PnP: H x 18 (rvec, tvec, R'(9), pad)
  int32_t* d_counts = nullptr;    // F: H x 3 ; PnP: H
  int* d_sel = nullptr;           // [0]=best flat index, [1]=n_iters, [2]=best count, [3]=n_records
  int* h_sel = nullptr;           // pinned
  double* d_pose = nullptr;       // 16 doubles: refined rvec,tvec + diagnostics
  double* h_pose = nullptr;       // pinned
  double* d_cam = nullptr;        // P1 (12), P2 (12), M (12) scratch
  // last-call debug views
  int last_pnp_h = 0, last_f_h = 0;

  // interleaved BGR staging (channels == 3): the image lands here, split_planes_kernel de-interleaves it
  uint8_t* d_bgr = nullptr;
  uint8_t* d_gray = nullptr;      // vo_bgr_to_gray staging (lazily allocated)
  void* d_sor = nullptr;          // SORcloud scratch: sort keys / order / sorted points / CUB temp (lazily allocated)
  size_t sor_tmp_bytes = 0;
  void* sgbm = nullptr;           // vo::Sgbm (sgbm.cu): dense-stereo buffers, allocated by the first SGBM call
  void* orb = nullptr;            // vo::Orb (orb.cu): ORB descriptor-stage buffers
  float* h_pts = nullptr;         // cap * 8 floats

  // LK work counters
  unsigned long long* d_lk_work = nullptr;  // [0]=point-levels, [1]=iterations
  unsigned long long* h_lk_work = nullptr;
};

namespace vo {

void set_error(const char* fmt, ...);

#define VO_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (call);                                                               \
    if (_e != cudaSuccess) {                                                               \
      vo::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));  \
      return VO_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

#define VO_TRY(expr)            \
  do {                          \
    int _r = (expr);            \
    if (_r != VO_OK) return _r; \
  } while (0)

// Launch bracket: counts launches, optionally records events around the launch.
struct LaunchScope {
  vo_ctx* c;
  int kind;
  cudaEvent_t a = nullptr, b = nullptr;
  LaunchScope(vo_ctx* ctx, int k);
  ~LaunchScope();
};

// ---------------------------------------------------------------------------- ordered compaction core
// Single pass with decoupled look-back (points.cu compact_kernel, ransac.cu mask_compact_kernel): every CTA owns a
// tile of CP_TILE elements (CP_ITEMS consecutive ones per thread), scans its per-thread counts, publishes its tile
// total tagged with the launch epoch, and sums the totals of the lower-indexed tiles (all co-resident, so the wait
// is short and cannot deadlock).  Returns the output position of the calling thread's first kept element; the last
// CTA writes the overall count and bumps the device-side epoch (graph-replay safe).
constexpr int CP_THREADS = 256;
constexpr int CP_ITEMS = 8;
constexpr int CP_TILE = CP_THREADS * CP_ITEMS;

#ifdef __CUDACC__
__device__ __forceinline__ int compact_tile_offset(int cnt, int* __restrict__ count_out, volatile unsigned long long* tile_state,
                                                   unsigned* epoch_ctr) {
  const unsigned epoch = *reinterpret_cast<volatile unsigned*>(epoch_ctr);
  __shared__ int warp_tot[CP_THREADS / 32];
  __shared__ int s_base;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  int incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  if (lane == 31) warp_tot[w] = incl;
  __syncthreads();
  if (w == 0) {
    // exclusive scan of the 8 warp totals, publish the tile total, look back
    int wt = lane < CP_THREADS / 32 ? warp_tot[lane] : 0;
    int wi = wt;
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += v;
    }
    const int tile_total = __shfl_sync(0xffffffffu, wi, CP_THREADS / 32 - 1);
    if (lane < CP_THREADS / 32) warp_tot[lane] = wi - wt;
    if (lane == 0) {
      tile_state[blockIdx.x] = ((unsigned long long)epoch << 32) | (unsigned)tile_total;
      __threadfence();
    }
    int base = 0;
    for (int j = lane; j < (int)blockIdx.x; j += 32) {
      unsigned long long v;
      do {
        v = tile_state[j];
      } while ((unsigned)(v >> 32) != epoch);
      base += (int)(unsigned)v;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) base += __shfl_xor_sync(0xffffffffu, base, d);
    if (lane == 0) {
      s_base = base;
      if (blockIdx.x == gridDim.x - 1) {
        *count_out = base + tile_total;
        *epoch_ctr = epoch + 1;
      }
    }
  }
  __syncthreads();
  return s_base + warp_tot[w] + incl - cnt;
}
#endif

constexpr size_t RES_BYTES = 16 * sizeof(double) + 32 * sizeof(int);

static inline int div_up(int a, int b) { return (a + b - 1) / b; }

// stage launchers (each returns VO_OK / VO_ERR_CUDA) -------------------------------------
int pyr_alloc(vo_ctx* c, Pyramid& p);
void pyr_free(Pyramid& p);
int pyr_build(vo_ctx* c, int slot, const uint8_t* d_tight /*w*h device*/, bool with_deriv);
int pyr_ensure_deriv(vo_ctx* c, int slot);
int pyr_split_bgr(vo_ctx* c, int slot, const uint8_t* d_bgr /*tight h x 3w*/);
int pyr_unpack_rows(vo_ctx* c, int slot, const uint8_t* d_src, int src_pitch);
int bgr2gray_launch(vo_ctx* c, const uint8_t* d_bgr, int src_pitch, uint8_t* d_gray, int dst_pitch);
int bgr2gray_launch_wh(vo_ctx* c, const uint8_t* d_bgr, int src_pitch, uint8_t* d_gray, int dst_pitch, int w, int h);
void sgbm_free(vo_ctx* c);
void orb_free(vo_ctx* c);
PyrView pyr_view(const Pyramid& p);

int lk_launch(vo_ctx* c, int slot_prev, int slot_next, const float2* d_prev, int n, float2* d_next, uint8_t* d_status,
              float* d_err);

int grid_launch(vo_ctx* c, int rows, int cols, int step, float2* d_xy, int* n_out);
// order-preserving compaction of up to three parallel arrays by flag==1; count -> d_count[slot]
int compact_launch(vo_ctx* c, const uint8_t* d_flags, int n, const float2* a_in, float2* a_out, const float2* b_in,
                   float2* b_out, const float3* c_in, float3* c_out, int32_t* idx_out, int count_slot);
// tracks of the points idx[0..n) gathered from another launch's outputs (look-ahead LK): out[j] = in[idx[j]]
int gather_tracks_launch(vo_ctx* c, const int32_t* d_idx, int n, const float2* trk_in, const uint8_t* st_in, float2* trk_out,
                         uint8_t* st_out);
int triangulate_launch(vo_ctx* c, const double* d_P1P2, const float2* a, const float2* b, int n, float3* out,
                       const double* d_M /*nullable: fused rigid transform -> out2*/, float3* out2);
int transform_launch(vo_ctx* c, const double* d_M, const float3* in, int n, float3* out);
int keyframe_epilogue_launch(vo_ctx* c, const double* pose3x4, const float3* cam, const float2* xy_in, int n, float3* world,
                             float2* xy_out);

int fmat_solve_launch(vo_ctx* c, const float2* m1, const float2* m2, const int32_t* d_samples, int h, double* d_models,
                      int32_t* d_counts);
int fmat_score_launch(vo_ctx* c, const float2* m1, const float2* m2, int n, const double* d_models, int32_t* d_counts,
                      int h, float thr2);
int fmat_mask_launch(vo_ctx* c, const float2* m1, const float2* m2, int n, const double* d_models, const int* d_sel,
                     float thr2, uint8_t* d_mask);
int pnp_solve_launch(vo_ctx* c, const float3* xyz, const float2* xy, const int32_t* d_samples, int h, double* d_models,
                     int32_t* d_counts);
int pnp_score_launch(vo_ctx* c, const float3* xyz, const float2* xy, int n, const double* d_models, int32_t* d_counts,
                     int h, float thr2);
int pnp_mask_launch(vo_ctx* c, const float3* xyz, const float2* xy, int n, const double* d_models, const int* d_sel,
                    float thr2, uint8_t* d_mask);
// solvePnPRansac with 4 or 5 points: the direct solve (P3P / EPnP on all points), every point an inlier, no refinement
int pnp_direct_launch(vo_ctx* c, const float3* xyz, const float2* xy, int n, int32_t* d_samples, double* d_models,
                      int32_t* d_counts, int* d_sel, int32_t* d_idx, int* d_n_inl, double* d_pose);
// findFundamentalMat's small-N estimators: LMedS for 8..14 points, the raw 7-point result for 7
int fmat_lmeds_launch(vo_ctx* c, const float2* m1, const float2* m2, int n, const double* d_models, const int32_t* d_counts,
                      int n_models, int* d_sel, uint8_t* d_mask);
int fmat_seven_launch(vo_ctx* c, const int32_t* d_counts, int n, int* d_sel, uint8_t* d_mask);
// fused-chain forms: scoring with the acceptance replay in its last CTA; best-model mask + ordered compaction in one launch
int fmat_score_select_launch(vo_ctx* c, const float2* m1, const float2* m2, int n, const double* d_models, int32_t* d_counts,
                             int h, float thr2, double conf, int max_iters, int* d_sel);
int fmat_mask_compact_launch(vo_ctx* c, const float2* m1, const float2* m2, const float3* xyz, int n, const double* d_models,
                             const int* d_sel, float thr2, uint8_t* d_mask, float2* o1, float2* o2, float3* oxyz,
                             int count_slot);
int pnp_score_select_launch(vo_ctx* c, const float3* xyz, const float2* xy, int n, const double* d_models, int32_t* d_counts,
                            int h, float thr2, double conf, int max_iters, int* d_sel);
int pnp_mask_compact_launch(vo_ctx* c, const float3* xyz, const float2* xy, int n, const double* d_models, const int* d_sel,
                            float thr2, uint8_t* d_mask, int32_t* d_idx, int count_slot);
// RANSAC record-setter scan: counts[n_models] (flattened, models_per_sample each) -> d_sel
constexpr int RNG_LEN = 1 << 17;
// device-side minimal-sample generation (OpenCV getSubset semantics); M = 7 checks collinearity
int sample_launch(vo_ctx* c, int model_points, const float2* m1, const float2* m2, int n_max, int h, int32_t* d_samples,
                  int* d_flag);
int select_launch(vo_ctx* c, const int32_t* d_counts, int n_samples, int models_per_sample, int model_points, int n_points,
                  double conf, int max_iters, int* d_sel);
int refine_init();   // constant tables of the refinement kernel (once per process)
int pnp_refine_launch(vo_ctx* c, const float3* xyz, const float2* xy, const int32_t* d_idx, const int* d_n_inl,
                      const double* d_models, const int* d_sel, double* d_pose);

int sor_mean_knn_launch(vo_ctx* c, const float3* d_pts, int n, int mean_k, float* d_mean);
int selfcheck_run(vo_ctx* c);
int epnp_debug_launch(vo_ctx* c, const float* d_obj, const float* d_img, double* d_dbg);
int anms_launch(vo_ctx* c, const float* h_xy, const float* h_resp, int n, int num_keep, int32_t* keep_idx, int cap,
                int* n_keep);
int synth_launch(vo_ctx* c, int seed, int frame, int eye, uint8_t* d_out);
int fp32_peak_launch(vo_ctx* c, double* tflops);
int int32_peak_launch(vo_ctx* c, double* tops);

}  // namespace vo
