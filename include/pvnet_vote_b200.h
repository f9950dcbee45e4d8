/*
 * pvnet_vote_b200.h -- C ABI of the B200-native RANSAC voting layer.
 *
 * Drop-in boundary for clean-pvnet's `lib/csrc/ransac_voting` (reference paths
 * below are relative to the clean-pvnet checkout):
 *
 *   reference interface                                         replaced by
 *   ----------------------------------------------------------  ---------------------------------
 *   ransac_voting_gpu.py:112  ransac_voting_layer_v3(...)        pvb_ransac_voting_v3
 *   ransac_voting_gpu.py:6    ransac_voting_layer(...)           pvb_ransac_voting_v3 (same result,
 *                                                                 see DESIGN.md "v1 vs v3")
 *   ransac_voting_gpu.py:202  estimate_voting_distribution_...   pvb_estimate_voting_distribution
 *   src/ransac_voting.cpp:20  generate_hypothesis                pvb_generate_hypothesis
 *   src/ransac_voting.cpp:41  voting_for_hypothesis              pvb_voting_for_hypothesis
 *   src/ransac_voting.cpp:64  generate_hypothesis_vanishing_pt   pvb_generate_hypothesis_vanishing_point
 *   src/ransac_voting.cpp:85  voting_for_hypothesis_vanishing_pt pvb_voting_for_hypothesis_vanishing_point
 *
 *   and, for the callers either side of the layer:
 *   lib/networks/pvnet/resnet18.py:65-76  decode_keypoint         pvb_decode_v3 (+ pvb_estimate_voting_distribution)
 *   lib/evaluators/linemod/pvnet.py:118-130  weight loop          pvb_uncertainty_weights
 *   lib/csrc/uncertainty_pnp/src/ext.h  uncertainty_pnp(...)      pvb_uncertainty_pnp (batched)
 *   un_pnp_utils.py:25-31  cv2.solvePnP(..., SOLVEPNP_P3P)        pvb_uncertainty_pnp_init
 *   evaluators/linemod/pvnet.py:118-130 + un_pnp_utils.py:6-57     pvb_uncertainty_pnp_from_votes (all three, one launch)
 *
 * Conventions
 *   - plain C: device pointers, sizes, strides (in ELEMENTS), a CUDA stream
 *     handle.  No torch types.  All work is enqueued on `stream`; no entry
 *     point synchronises or allocates (the caller owns the workspace), so the
 *     calls are CUDA-graph capturable.  Exceptions: the *_host variants, which
 *     take HOST buffers and run their own copy/compute pipeline, and the setup
 *     calls of pvb_exchange (create / connect / destroy).
 *   - every function returns PVB_OK (0) or a pvb_status error code;
 *     pvb_last_error() returns a thread-local description.  Nothing calls
 *     exit()/abort() (the reference's gpuErrchk does, cuda_common.h:19-25).
 *   - there is NO CPU implementation behind this ABI.  If no CUDA device is
 *     usable the calls fail with PVB_ERR_CUDA.
 */
#ifndef PVNET_VOTE_B200_H_
#define PVNET_VOTE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVB_VERSION 200

#if defined(__GNUC__)
#define PVB_API __attribute__((visibility("default")))
#else
#define PVB_API
#endif

typedef void *pvb_stream_t; /* cudaStream_t / CUstream, 0 = legacy default stream */

typedef enum pvb_status {
    PVB_OK = 0,
    PVB_ERR_INVALID = 1,   /* bad argument (null pointer, shape, dtype, stride) */
    PVB_ERR_CUDA = 2,      /* CUDA runtime / launch failure, or no usable device */
    PVB_ERR_WORKSPACE = 3, /* workspace too small or misaligned */
    PVB_ERR_CAPACITY = 4,  /* more pixels selected than `capacity` (reported by pvb_read_status) */
    PVB_ERR_TIMEOUT = 5    /* pvb_exchange_wait gave up: a rank never published its result */
} pvb_status;

/* element type of the mask tensor (reference: any dtype goes through .byte(), ransac_voting_gpu.py:125) */
typedef enum pvb_mask_dtype {
    PVB_MASK_U8 = 0, /* also torch.bool */
    PVB_MASK_I8 = 1,
    PVB_MASK_I16 = 2,
    PVB_MASK_I32 = 3,
    PVB_MASK_I64 = 4, /* torch.argmax output, resnet18.py:69 */
    PVB_MASK_F32 = 5,
    PVB_MASK_F64 = 6
} pvb_mask_dtype;

/* how foreground pixels are picked */
typedef enum pvb_select_mode {
    PVB_SELECT_BYTE = 0, /* v3: (uint8)mask != 0, fg = sum of byte values   (ransac_voting_gpu.py:125-126) */
    PVB_SELECT_EQ1 = 1   /* distribution: mask == 1, fg = pixel count       (ransac_voting_gpu.py:207-208) */
} pvb_select_mode;

/* Problem descriptor shared by the layer-level entry points. */
typedef struct pvb_desc {
    int32_t B, H, W, K;       /* batch, image height/width, keypoints (vn) */
    int32_t hn;               /* hypotheses per (image, keypoint): round_hyp_num for v3,
                                 round_hyp_num*ceil(min_hyp_num/round_hyp_num) for the distribution */
    float inlier_thresh;      /* cos threshold, compared as fp32 like the reference (.cu:124) */
    int32_t min_num, max_num; /* ransac_voting_gpu.py:129,135 */
    int32_t mask_dtype;       /* pvb_mask_dtype */
    int32_t select_mode;      /* pvb_select_mode */
    int64_t mask_stride[3];   /* strides of mask [B,H,W], elements */
    int64_t vertex_stride[5]; /* strides of vertex [B,H,W,K,2], elements (any layout, e.g. the permuted
                                 NCHW view decode_keypoint passes, resnet18.py:66-68) */
    int32_t capacity;         /* max selected pixels per image held by the workspace; 0 = default
                                 (min(H*W, max_num + 8*sqrt(max_num) + 64)); H*W is always safe */
    int32_t img_base;         /* global index of image 0 (multi-GPU shards keep one philox stream) */
    uint64_t seed;            /* philox seed, used when idxs / selection are NULL */
    int32_t rng_tag_idx;      /* philox stream tags (see DESIGN.md "Sampling"); 0 = per-op default */
    int32_t rng_tag_sel;
} pvb_desc;

/* Offsets (bytes from the workspace base) of the intermediate buffers; for tests and tooling. */
typedef struct pvb_layout {
    size_t total;    /* == pvb_workspace_bytes() */
    size_t status;   /* int32[4]: [0] sticky error code (pvb_status), [1] image that overflowed */
    size_t fgsum;    /* uint64[B]  sum of mask bytes (BYTE) / count (EQ1) */
    size_t nz;       /* int32[B]   selected-before-thinning count */
    size_t tn;       /* int32[B]   selected pixel count after thinning (0 when skipped) */
    size_t state;    /* int32[B]   0 ok, 1 skipped (fg < min_num) */
    size_t bits;     /* uint32[B][nwords] selection bitmap (after thinning), bit j of word w = pixel 32*w+j */
    size_t ticket;   /* int32[B] arrival counter of the compaction CTAs (block ids are drawn in arrival order) */
    size_t blocktot; /* uint32[B][nblocks] selected pixels per 128-word block | bit 31 "published" (decoupled look-back) */
    size_t xy;       /* float2[B][capacity]   (x,y) of the t-th selected pixel, row-major (torch.nonzero) order */
    size_t dirs;     /* float2[B][K][capacity] gathered vertex vectors (k-major) */
    size_t hyp;      /* float2[B][K][hn] */
    size_t counts;   /* int32[B][K][hn] */
    size_t win;      /* float2[B][K] winning hypothesis before the refit */
    size_t refit_partial; /* double[B][K][refit_splits][5] partial normal equations */
    size_t refit_ticket;  /* int32[B][K] arrival counters of the refit CTAs */
    int32_t nwords;  /* ceil(H*W/32) */
    int32_t nblocks; /* ceil(nwords/128) */
    int32_t capacity;
    int32_t refit_splits;
} pvb_layout;

PVB_API int pvb_version(void);
PVB_API const char *pvb_last_error(void);

/* Workspace sizing.  The workspace must be 256-byte aligned device memory. */
PVB_API size_t pvb_workspace_bytes(const pvb_desc *d);
PVB_API int pvb_workspace_layout(const pvb_desc *d, pvb_layout *out);

/* ransac_voting_layer_v3 (ransac_voting_gpu.py:112-199), whole batch, no host sync.
 *   mask    device, [B,H,W] of d->mask_dtype with d->mask_stride
 *   vertex  device fp32, [B,H,W,K,2] with d->vertex_stride
 *   idxs    optional device int32 [B,hn,K,2] contiguous: the reference's per-image `idxs`
 *           (:145).  NULL -> drawn in-kernel from the philox stream (seed, tag, image, k, h).
 *   selection optional device fp32 [B,H,W] contiguous: the reference's U(0,1) `selection`
 *           (:136), consulted only for images with fg > max_num.  NULL -> philox.
 *   out_kpt device fp32 [B,K,2] contiguous.
 * `confidence` / `max_iter` of the reference do not influence its result (idxs is drawn once,
 * outside the loop, :145 vs :150) and therefore have no counterpart here. */
PVB_API int pvb_ransac_voting_v3(const pvb_desc *d, const void *mask, const float *vertex,
                         const int32_t *idxs, const float *selection, float *out_kpt,
                         void *workspace, size_t workspace_bytes, pvb_stream_t stream);

/* Fused front end of Resnet18.decode_keypoint (lib/networks/pvnet/resnet18.py:65-76): same as
 * pvb_ransac_voting_v3, but the mask is torch.argmax(seg, 1) (:69) computed on the fly from the fp32 logits
 *   seg  device fp32 [B,classes,H,W]; d->mask_stride = its (B,H,W) strides, class_stride its class stride (elements);
 *        d->mask_dtype is ignored.  First maximal class wins, NaN counts as maximal (torch.argmax semantics).
 *   mask_out optional device int64 [B,H,W] contiguous: receives the argmax mask decode_keypoint returns (:73,:76).
 * d->select_mode applies to the class index exactly as it would to the mask tensor. */
PVB_API int pvb_decode_v3(const pvb_desc *d, const float *seg, int32_t classes, int64_t class_stride, int64_t *mask_out,
                          const float *vertex, const int32_t *idxs, const float *selection, float *out_kpt,
                          void *workspace, size_t workspace_bytes, pvb_stream_t stream);

/* estimate_voting_distribution_with_mean (ransac_voting_gpu.py:202-274).
 *   mean device fp32 [B,K,2];  out_cov device fp32 [B,K,2,2].  d->select_mode must be
 *   PVB_SELECT_EQ1 to match the reference (:207).  idxs optional int32 [B,hn,K,2] (the 16
 *   per-round draws of :235 concatenated in round order). */
PVB_API int pvb_estimate_voting_distribution(const pvb_desc *d, const void *mask, const float *vertex,
                                     const float *mean, const int32_t *idxs, const float *selection,
                                     float *out_cov, void *workspace, size_t workspace_bytes,
                                     pvb_stream_t stream);

/* Weights of the uncertainty PnP (lib/evaluators/linemod/pvnet.py:118-130, a scipy.linalg.sqrtm + np.linalg.inv loop
 * per keypoint on the CPU): weights[i] = (wxx, wxy, wyy) of inv(sqrtm(cov[i])), zeros where cov[i][0][0] < 1e-6, any
 * entry is NaN or cov[i] is not positive definite.  cov device fp32 [n,2,2] (16-byte aligned), weights device fp32 [n,3].
 * The result is what un_pnp_utils.uncertainty_pnp (lib/csrc/uncertainty_pnp/un_pnp_utils.py:6) takes as weights_2d. */
PVB_API int pvb_uncertainty_weights(const float *cov, float *weights, int32_t n, pvb_stream_t stream);

/* Batched twin of the reference's C entry `uncertainty_pnp(pts2d, pts3d, wgt2d, K, init_rt, result_rt, pn)`
 * (lib/csrc/uncertainty_pnp/src/ext.h:1-9, uncertainty_pnp.cpp:61-92; bound through cffi by un_pnp_utils.py:49-53): refines
 * n poses (angle-axis + translation, 6 doubles) by minimising the weighted reprojection error of pn points each with the
 * Levenberg-Marquardt trust-region loop and the default options of Ceres Solver 2.0 (the reference's minimiser; pinned
 * against the reference's own Ceres binary: tests/golden/ceres_pnp.npz).  One warp per problem, fp64
 * like the reference.
 * All pointers are DEVICE memory: pts2d [n,pn,2], wgt2d [n,pn,3] = (wxx,wxy,wyy), init_rt / result_rt [n,6],
 * pts3d [pn,3] and K [3,3] (row-major) per problem at pts3d + p*pts3d_stride / K + p*k_stride (strides in doubles; 0 = one
 * array shared by all problems), info optional int32 [n,2] = (iterations, termination: 1 gradient, 2 parameter, 3 function
 * tolerance, 4 trust region collapsed, 5 iteration limit, 6 five invalid steps in a row).  options NULL = Ceres defaults. */
typedef struct pvb_pnp_options {
    int32_t max_num_iterations;       /* 50   */
    int32_t reserved;                 /* 0    */
    double function_tolerance;        /* 1e-6 */
    double gradient_tolerance;        /* 1e-10 */
    double parameter_tolerance;       /* 1e-8 */
} pvb_pnp_options;
PVB_API int pvb_uncertainty_pnp(const double *pts2d, const double *pts3d, const double *wgt2d, const double *K,
                                const double *init_rt, double *result_rt, int32_t *info, int32_t n, int32_t pn,
                                int64_t pts3d_stride, int64_t k_stride, const pvb_pnp_options *options, pvb_stream_t stream);

/* The whole un_pnp tail of the evaluator in ONE launch, straight from the voting layer's fp32 outputs:
 *   lib/evaluators/linemod/pvnet.py:118-130  weights = inv(sqrtm(var)) per keypoint          (pvb_uncertainty_weights)
 *   un_pnp_utils.py:25-31                    P3P initial pose on the 4 best-weighted points  (pvb_uncertainty_pnp_init)
 *   un_pnp_utils.py:49-53 -> ext.h           Ceres refinement                                (pvb_uncertainty_pnp)
 * kpt_2d device fp32 [n,pn,2]; exactly one of cov (device fp32 [n,pn,2,2], 16-byte aligned) and weights (device fp32
 * [n,pn,3]); pts3d / K as in pvb_uncertainty_pnp; init_rt optional [n,6] (NULL: P3P); result_rt [n,6]; optional outputs:
 * init_out [n,6] (the initial pose used), weights_out fp32 [n,pn,3], info [n,2].  pn <= 64.  Bit-identical to running the
 * three entry points one after the other on the same data (tests/test_gpu_pnp.py). */
PVB_API int pvb_uncertainty_pnp_from_votes(const float *kpt_2d, const float *cov, const float *weights, const double *pts3d,
                                           const double *K, const double *init_rt, double *result_rt, double *init_out,
                                           float *weights_out, int32_t *info, int32_t n, int32_t pn, int64_t pts3d_stride,
                                           int64_t k_stride, const pvb_pnp_options *options, pvb_stream_t stream);

/* Initial poses for pvb_uncertainty_pnp, the reference's recipe on the device (un_pnp_utils.py:25-31:
 * `idxs = argsort(wxx + wxy)[-4:]`, `cv2.solvePnP(points_3d[idxs], points_2d[idxs], K, ..., flags=cv2.SOLVEPNP_P3P)`): P3P on
 * the 2nd..4th best-weighted keypoints, the best-weighted one chooses among the (up to four) poses by its reprojection
 * error.  Same layouts as pvb_uncertainty_pnp; writes init_rt [n,6] (angle-axis, translation); a problem without an
 * admissible solution gets NaNs (what OpenCV returns there).  pn >= 4.  The arithmetic is pinned against cv2.solvePnP on
 * the CPU (tests/test_p3p_host_core.py), the device launch against OpenCV on the GPU box (tests/test_gpu_zz_p3p.py). */
PVB_API int pvb_uncertainty_pnp_init(const double *pts2d, const double *pts3d, const double *wgt2d, const double *K,
                                     double *init_rt, int32_t n, int32_t pn, int64_t pts3d_stride, int64_t k_stride,
                                     pvb_stream_t stream);

/* Reads the sticky status word of a workspace (synchronises `stream`). */
PVB_API int pvb_read_status(const pvb_desc *d, const void *workspace, pvb_stream_t stream);

/* HOST-buffer variant of pvb_ransac_voting_v3: mask/vertex/out_kpt are host pointers
 * (pinned for full speed), contiguous [B,H,W] / [B,H,W,K,2] / [B,K,2].  Splits the batch into
 * `chunk_images`-sized pieces on three internal streams so that one piece's PCIe traffic overlaps the others'
 * kernels.  What crosses the bus is chosen per call by `flags`:
 *   0 (default)            the mask is staged -- one contiguous cudaMemcpyAsync per piece, the copy engine's PCIe
 *                          rate -- and the vertex field is read IN PLACE from the pinned host tensor: gather fetches only the
 *                          SELECTED pixels' rows (tn*K*8 bytes per image instead of the dense H*W*K*8)
 *   PVB_HOST_STAGE_VERTEX  also copy the dense vertex field (pageable inputs are always staged)
 *   PVB_HOST_INPLACE_MASK  read the mask in place as well (no DMA at all; round 1's mode)
 * Stream-ordered after the work already queued on `stream`; `stream` is synchronised before the
 * call returns (the result is in host memory), so CUDA events recorded on `stream` around the call
 * bracket all copies and kernels.  dev_scratch: 256-byte aligned device memory of
 * pvb_host_scratch_bytes(d, chunk_images) bytes.  Thread-safe (per-thread, per-device streams). */
#define PVB_HOST_STAGE_VERTEX 2u
#define PVB_HOST_INPLACE_MASK 4u
PVB_API size_t pvb_host_scratch_bytes(const pvb_desc *d, int32_t chunk_images);
PVB_API int pvb_ransac_voting_v3_host(const pvb_desc *d, const void *mask_host, const float *vertex_host,
                              float *out_kpt_host, int32_t chunk_images, uint32_t flags,
                              void *dev_scratch, size_t dev_scratch_bytes, pvb_stream_t stream);

/* ---- multi-GPU: images are sharded across ranks, one process per GPU --------------------------------
 * The path has no exchange inside the algorithm; what crosses GPUs is each rank's [B_r,K,2] keypoints becoming visible on
 * every rank.  The reference has no counterpart (torch.nn.DataParallel in the trainer only, lib/train/trainers/trainer.py:11).
 * A pvb_exchange is a receive ring in this rank's HBM -- recv[slots][world][floats_per_rank] of 8-byte words
 * {float bits, seq} -- mapped into every peer through CUDA IPC.  pvb_ransac_voting_v3_push is pvb_ransac_voting_v3 whose
 * refit kernel ALSO stores every (image, keypoint) result, as it is produced, into slot (seq-1)%slots of every peer's ring
 * over NVLink: aligned 8-byte stores are single-copy atomic, so every word validates itself and the producer needs no
 * fence, counter or flag and never waits.  pvb_exchange_wait enqueues a one-CTA kernel that polls this rank's own ring
 * until every expected word carries `seq` (bounded by timeout_s) and writes the floats to `out` (device fp32
 * [world][bytes_per_rank/4]; floats_per_rank: HOST array of world counts for ragged shards, NULL = every rank publishes
 * bytes_per_rank/4 floats).  Ring discipline (caller): before call `seq` is launched, the wait of call `seq - slots/2` must
 * already be enqueued on the same stream on every rank; seq starts at 1 and increases by 1 per call on every rank alike.
 * Setup (once): create -> get_handle -> all-gather the 64-byte handles by any means (e.g.
 * torch.distributed.all_gather_object) -> connect.  connect_ptrs takes raw base pointers instead (ranks that live in one
 * process, or memory mapped by other means). */
typedef struct pvb_exchange pvb_exchange;
#define PVB_IPC_HANDLE_BYTES 64
PVB_API int pvb_exchange_create(int32_t rank, int32_t world, int32_t slots, size_t bytes_per_rank, pvb_exchange **out);
PVB_API size_t pvb_exchange_bytes_per_rank(const pvb_exchange *ex); /* bytes_per_rank rounded up to 16 */
PVB_API void *pvb_exchange_base(const pvb_exchange *ex);
PVB_API int pvb_exchange_get_handle(const pvb_exchange *ex, void *handle /* PVB_IPC_HANDLE_BYTES */);
PVB_API int pvb_exchange_connect(pvb_exchange *ex, const void *handles /* world * PVB_IPC_HANDLE_BYTES, rank order */);
PVB_API int pvb_exchange_connect_ptrs(pvb_exchange *ex, void *const *bases /* world base pointers, rank order */);
PVB_API int pvb_ransac_voting_v3_push(const pvb_desc *d, const void *mask, const float *vertex, const int32_t *idxs,
                                      const float *selection, float *out_kpt, void *workspace, size_t workspace_bytes,
                                      pvb_exchange *exchange, uint64_t seq, pvb_stream_t stream);
/* the same for the second half of the un_pnp pair (resnet18.py:71-72): every [2,2] covariance goes to every peer as the
 * covariance kernel produces it (4 floats per (image, keypoint); use an exchange of its own: bytes_per_rank >= B*K*16) */
PVB_API int pvb_estimate_voting_distribution_push(const pvb_desc *d, const void *mask, const float *vertex, const float *mean,
                                                  const int32_t *idxs, const float *selection, float *out_cov, void *workspace,
                                                  size_t workspace_bytes, pvb_exchange *exchange, uint64_t seq,
                                                  pvb_stream_t stream);
PVB_API int pvb_exchange_wait(pvb_exchange *ex, uint64_t seq, void *out, const int32_t *floats_per_rank, double timeout_s,
                              pvb_stream_t stream);
PVB_API int pvb_exchange_status(pvb_exchange *ex, pvb_stream_t stream); /* synchronises; PVB_ERR_TIMEOUT after a timed-out wait */
PVB_API int pvb_exchange_destroy(pvb_exchange *ex);

/* Stage timing (tooling, used by bench.py).  pvb_profile_enable(n): n = 0 off; n >= 1: every n-th call of a layer entry
 * point records 5 CUDA events on the launching stream at its stage boundaries (an event record drains the pipeline between
 * two kernels, ~3 us each on B200: sample with n > 1 to keep the profile out of a measurement).  pvb_profile_read()
 * synchronises those events and ADDS the elapsed milliseconds of every profiled call since the last pvb_profile_reset()
 * into ms[PVB_STAGE_COUNT], returning the number of calls accumulated.  Per host thread. */
enum { PVB_STAGE_SELECT = 0,   /* mask_bits + thin_gather (thinning, ordered compaction, vertex gather) */
       PVB_STAGE_GENERATE = 1, /* hypothesis generation */
       PVB_STAGE_VOTE = 2,     /* counts memset + vote kernel (the dominant kernel) */
       PVB_STAGE_FINISH = 3,   /* winner + refit, or covariance */
       PVB_STAGE_COUNT = 4 };
PVB_API int pvb_profile_enable(int32_t every);
/* Tuning switches (tooling for A/B measurements; process-wide, atomic).  Results do not depend on them.
 *   gather_mode  access pattern of the gather kernel on an interleaved vertex tensor in device memory: 0 = auto = 1 =
 *                pixel-wise (one lane per pixel), 2 = row-wise (a warp reads whole 8*K-byte pixel rows; what in-place
 *                host reads always use)
 *   vote_variant pixel tile of the vote kernel: 0 = 1 = 512 pixels (default), 2 = 256, 3 = 1024 */
PVB_API int pvb_set_tuning(int32_t gather_mode, int32_t vote_variant);
PVB_API int pvb_profile_reset(void);
PVB_API int pvb_profile_read(double *ms, int32_t n);

/* ---- twins of the reference pybind module `ransac_voting` (ransac_voting.cpp:102-107) ---- */
/* direct [tn,vn,2] f32, coords [tn,2] f32, idxs [hn,vn,2] i32 -> hyp [hn,vn,2] f32 (fully written;
 * degenerate pairs give (0,0), the reference's zero fill, .cu:42-43,75) */
PVB_API int pvb_generate_hypothesis(const float *direct, const float *coords, const int32_t *idxs, float *hyp,
                            int32_t tn, int32_t vn, int32_t hn, pvb_stream_t stream);
/* sets inliers[h,k,t]=1 (uint8 [hn,vn,tn]) where the test passes; other bytes untouched (.cu:124-125) */
PVB_API int pvb_voting_for_hypothesis(const float *direct, const float *coords, const float *hyp, uint8_t *inliers,
                              int32_t tn, int32_t vn, int32_t hn, float inlier_thresh, pvb_stream_t stream);
/* hyp [hn,vn,3] */
PVB_API int pvb_generate_hypothesis_vanishing_point(const float *direct, const float *coords, const int32_t *idxs,
                                            float *hyp, int32_t tn, int32_t vn, int32_t hn,
                                            pvb_stream_t stream);
PVB_API int pvb_voting_for_hypothesis_vanishing_point(const float *direct, const float *coords, const float *hyp,
                                              uint8_t *inliers, int32_t tn, int32_t vn, int32_t hn,
                                              float inlier_thresh, pvb_stream_t stream);

/* Fused count of the above (voting_for_hypothesis + torch.sum(dim 2), ransac_voting_gpu.py:156-159)
 * on the reference layouts: counts int32 [hn,vn].  Same kernel the layer uses. */
PVB_API int pvb_vote_count(const float *direct, const float *coords, const float *hyp, int32_t *counts,
                   int32_t tn, int32_t vn, int32_t hn, float inlier_thresh,
                   void *workspace, size_t workspace_bytes, pvb_stream_t stream);
PVB_API size_t pvb_vote_count_workspace_bytes(int32_t tn, int32_t vn, int32_t hn);

#ifdef __cplusplus
}
#endif
#endif /* PVNET_VOTE_B200_H_ */
