/*
 * isg.h — C ABI of libisg.so: the B200 (sm_100a) decode hot path of
 * aspirantll/instance-segmentation (utils/decode.py, utils/kmeans.py, utils/nms.py).
 *
 * The reference has no FFI: its boundary is a set of Python module functions
 * (SURVEY.md §8b).  Each entry point below names the reference function
 * (file:line, relative to the reference tree) whose arithmetic it replaces.
 *
 * Conventions (every function):
 *   - all pointers are DEVICE pointers unless the parameter is documented as
 *     "host"; the caller owns every buffer (inputs, outputs, workspace);
 *   - nothing is allocated, no global mutable state, nothing throws;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); functions
 *     do not synchronise unless documented ("blocking");
 *   - return value: 0 = ok, < 0 = ISG_E* (bad argument etc.), > 0 = cudaError_t;
 *   - image planes are row-major fp32; "plane stride"/"image stride" are in
 *     elements; rows are contiguous (stride W).
 *   - coordinates: pixel = (y, x) = (row, col) as in the reference's
 *     kp_mask.nonzero() (utils/decode.py:312); boxes are (x1, y1, x2, y2).
 */
#ifndef ISG_H_
#define ISG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISG_ABI_VERSION 6

#define ISG_OK            0
#define ISG_EINVAL       (-1)  /* bad argument (null pointer, negative extent, k > H*W ...) */
#define ISG_EWORKSPACE   (-2)  /* workspace too small or misaligned */
#define ISG_EUNSUPPORTED (-3)  /* size outside the supported range (see each function) */
#define ISG_ENOTCONVERGED (-4) /* isg_kmeans hit max_iter */

typedef void* isg_stream_t;    /* cudaStream_t */

/* words per seed record (32 B): {y0,y1,x0,x1 : int32 inclusive in-box bounds; cy,cx : fp32 grid
 * coordinate of the truncated box centre; 2 words padding} */
#define ISG_SEED_WORDS 8
/* words per ghost-filter record (16 B): {xlo,xhi,ylo,yhi : fp32, strict bounds} */
#define ISG_GHOST_WORDS 4
/* words per instance-statistics record: {count, ymin, xmin, ymax, xmax : int32} */
#define ISG_STAT_WORDS 5

/* NMS conventions */
#define ISG_NMS_PLUS1_LE 0  /* utils/nms.py:19,31-36: areas and overlaps use +1, survivor iff IoU <= thr (fp32 thr) */
#define ISG_NMS_TV_GT    1  /* torchvision nms per class on the ORIGINAL coordinates (_batched_nms_vanilla): no +1, suppress iff IoU > thr */
#define ISG_NMS_TV_TRICK 2  /* torchvision _batched_nms_coordinate_trick: before the IoU every box is shifted by
                             * class * (largest coordinate of the image's candidates + 1), all in fp32 - what
                             * batched_nms at utils/decode.py:400 does in torchvision 0.5.0 (the reference's pin) */
#define ISG_NMS_TV_BATCHED 3 /* batched_nms of current torchvision on CPU tensors: the coordinate trick for up to 1000
                             * candidates (boxes.numel() <= 4000), per-class NMS on the original coordinates above */
#define ISG_NMS_MAX_BOXES 16384

/* k-means metrics (utils/kmeans.py:34-39) */
#define ISG_KMEANS_EUCLIDEAN 0
#define ISG_KMEANS_COSINE    1

int         isg_abi_version(void);
const char* isg_strerror(int code);
/* Experiment hook (tests, tools/sweep_dense.py): the ISG_* tuning variables are read from the environment once, at the
 * first call into the library; this re-reads them.  Not for production use; not thread-safe against running calls. */
void        isg_debug_reload_tuning(void);
/* host helper: the device address of page-locked, mapped HOST memory (cudaHostAlloc / torch pin_memory()), for the
 * entry points that accept such a pointer in place of a device pointer (documented per parameter: the `ae` of
 * isg_assign_sparse and the `regression` of isg_decode_boxes are only read at the selected pixels / candidate anchors,
 * so they can stay in host memory and be gathered over PCIe).  Returns a cudaError_t (> 0) if the memory is not mapped. */
int         isg_host_device_pointer(const void* host_ptr, void** device_ptr);
/* host query: 1 if `device` is a compute-capability 10.x part this library was built for, else 0 */
int         isg_device_supported(int device);

/* ------------------------------------------------------------------------------------------
 * K2 — boundary-keypoint selection.  Replaces select_points (utils/decode.py:71-85) and
 * nms_hm (utils/decode.py:42-48).
 *   sel(p)  = kp[p] is among the k largest values of the image            (topk, :81)
 *   v(p)    = sel(p) ? kp[p] : 0                                          (mat*mask, :84)
 *   keep(p) = sel(p) && v(p) == max over the 3x3 window clipped to the image of v   (:45-47,85)
 * With ties at the k-th value every tied pixel is selected (torch.topk's choice is unspecified).
 * ------------------------------------------------------------------------------------------ */
size_t isg_topk_workspace_bytes(int B, int H, int W, int k);
/* k-th largest value per image as an order-preserving uint32 key (see isg_float_key in DESIGN.md).
 * kp: [B] images of H*W fp32, image b at kp + b*img_stride.  k in [0, H*W]; k > H*W -> ISG_EINVAL
 * (the reference's topk raises).  k == 0 selects nothing (key = 0xFFFFFFFF). */
int isg_topk_threshold(const float* kp, int B, int H, int W, int64_t img_stride, int k,
                       uint32_t* thr_key /*[B]*/, void* ws, size_t ws_bytes, isg_stream_t stream);
/* keep mask from a threshold.  keepbits: [B,H,ceil(W/32)] uint32, bit i of word w = pixel x = 32w+i.
 * mask_u8 (nullable): [B,H,W] uint8 0/1, the tensor select_points returns. */
int isg_keep_points(const float* kp, int B, int H, int W, int64_t img_stride, const uint32_t* thr_key,
                    uint32_t* keepbits, uint8_t* mask_u8, isg_stream_t stream);
/* isg_topk_threshold + isg_keep_points.  Workspaces must be 256-byte aligned. */
size_t isg_select_points_workspace_bytes(int B, int H, int W, int k);
int isg_select_points(const float* kp, int B, int H, int W, int64_t img_stride, int k,
                      uint32_t* keepbits, uint8_t* mask_u8, void* ws, size_t ws_bytes, isg_stream_t stream);
/* nms_hm (utils/decode.py:42-48): keep[p] = heat[p] == max over kernel x kernel window (stride 1,
 * -inf padding).  heat: [planes,H,W]; kernel odd, 1..15. */
int isg_nms_hm(const float* heat, int planes, int H, int W, int kernel, uint8_t* keep, isg_stream_t stream);
/* kp_mask.nonzero() (utils/decode.py:312): row-major (y,x) int32 pairs of the set bits.
 * idx: [B,cap,2]; count: [B] = number of set bits (may exceed cap; only the first cap are written). */
int isg_compact_points(const uint32_t* keepbits, int B, int H, int W, int cap,
                       int32_t* idx, int32_t* count, isg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Seeds.  Replaces decode_single's box -> centre/size arithmetic (utils/decode.py:428-432) and the
 * seed tensors of group_kp (:316-322) plus the per-instance ghost-filter bounds (:339-352).
 * rois: [B,Nmax,4] fp32, layout ISG_BOX_XYXY (x1,y1,x2,y2: the dict decode_boxes returns) or
 * ISG_BOX_CYCXHW (cy,cx,h,w: group_kp's center_indexes / center_whs arguments); n_seeds: [B] int32.
 * ys: [H], xs: [W] fp32 coordinate tables (utils/utils.py:453-458 sliced as at utils/decode.py:304).
 * ghost_k = fp32(0.5 + wh_delta), or < 0 to disable the ghost filter (every pixel passes);
 * scale = compute_scale() (utils/decode.py:34-35, 1 by default).
 * seeds: [B,Nmax,ISG_SEED_WORDS] 32-bit words; ghost: [B,Nmax,ISG_GHOST_WORDS] fp32.
 * ------------------------------------------------------------------------------------------ */
#define ISG_BOX_XYXY   0
#define ISG_BOX_CYCXHW 1
int isg_build_seeds(const float* rois, int layout, const int32_t* n_seeds, int B, int Nmax,
                    const float* ys, const float* xs, int H, int W, float ghost_k, float scale,
                    uint32_t* seeds, float* ghost, isg_stream_t stream);

/* isg_gather_kept + isg_build_seeds (XYXY) + isg_stats_init in one launch, for the batched pipeline: the kept
 * candidates of isg_box_nms become the detection tables (rois / scores / cls / n_out, as isg_gather_kept) and, in the
 * same pass, the seed records, ghost bounds and (stats nullable) reset statistics of the decode.  img_total (nullable):
 * the per-image point counters of isg_instance_polygons, zeroed here so that it can be called with totals_zeroed = 1. */
int isg_gather_build_seeds(const float* cand_boxes, const float* cand_scores, const int32_t* cand_cls,
                           const int32_t* keep, const int32_t* n_keep, int B, int cap, int Nmax,
                           const float* ys, const float* xs, int H, int W, float ghost_k, float scale,
                           float* rois, float* scores, int32_t* cls, int32_t* n_out,
                           uint32_t* seeds, float* ghost, int32_t* stats, int32_t* img_total, isg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K1+K3 — embedding + Gaussian membership + assignment.  Replaces group_kp's arithmetic core
 * (utils/decode.py:303-328) and the per-instance ghost filter / pixel count (:337-356).
 *   e = tanh(ae[0:2]) + grid ; s = exp(ae[2:4])           (per PIXEL)       (:305,315)
 *   P[m,j] = exp(-((e_y-c_jy)^2*s_y + (e_x-c_jx)^2*s_x)) * inbox[m,j]       (:325-327, no FMA)
 *   score, label = max_j P (first index on ties; all-zero row -> label 0)   (:328)
 * ae: image b at ae + b*img_stride, plane c at + c*plane_stride (4 planes).  The input is NOT
 * modified (the reference overwrites ae[0:2] in place, utils/decode.py:305; no caller reads it).
 * flag[m] = 1 iff the pixel passes its instance's strict ghost filter (:351-353).
 * stats (nullable): [B,Nmax,ISG_STAT_WORDS] int32, must be pre-initialised by isg_stats_init;
 * receives count / bbox of the flagged pixels of each instance.
 * ------------------------------------------------------------------------------------------ */
int isg_stats_init(int32_t* stats, int B, int Nmax, isg_stream_t stream);
/* sparse (reference-faithful): only the compacted keep pixels.  idx [B,cap,2], count [B] from
 * isg_compact_points.  label [B,cap] int32, score [B,cap] fp32, flag [B,cap] uint8. */
int isg_assign_sparse(const float* ae, int64_t img_stride, int64_t plane_stride,
                      const int32_t* idx, const int32_t* count, int cap,
                      const uint32_t* seeds, const float* ghost, const int32_t* n_seeds, int B, int Nmax,
                      int H, int W, const float* ys, const float* xs,
                      int32_t* label, float* score, uint8_t* flag, int32_t* stats, isg_stream_t stream);
/* label_map[b, idx[b,m]] = label[b,m] for the compacted keep pixels (m < min(count[b], cap)); every other element of
 * label_map is left untouched.  Lets the consumers of the dense label map (isg_instance_polygons, which looks labels up
 * at keep pixels only) run behind isg_assign_sparse. */
int isg_scatter_labels(const int32_t* idx, const int32_t* count, int cap, const int32_t* label, int B, int H, int W,
                       int32_t* label_map, isg_stream_t stream);
/* emb[b,m] = (e_y, e_x) = tanh(ae[0:2]) + grid at the compacted keep pixels (utils/decode.py:305,313-314): the points
 * the seeded k-means refinement of BASELINE config 4 clusters (X = e[M,2] for isg_kmeans).  emb: [B,cap,2] fp32. */
int isg_gather_embeddings(const float* ae, int64_t img_stride, int64_t plane_stride, const int32_t* idx,
                          const int32_t* count, int cap, int B, int H, int W, const float* ys, const float* xs,
                          float* emb, isg_stream_t stream);
/* dense fused: every pixel.  Reads kp (+1-pixel halo) and the 4 ae planes once, applies the
 * top-k threshold and the 3x3 peak test, assigns every pixel, writes label_map [B,H,W] int32,
 * keepbits [B,H,ceil(W/32)], optional score_map [B,H,W] fp32 (nullable), and accumulates stats
 * for the keep pixels.  label_map[keep] equals isg_assign_sparse's label (rows are independent). */
/* workspace: isg_assign_dense_workspace_bytes() bytes, 16-byte aligned, uninitialised: the tile scheduler words and
 * the per-tile seed lists in it are (re)written by every call (or by isg_build_tile_lists).  One workspace per
 * concurrently running call. */
size_t isg_assign_dense_workspace_bytes(int B, int Nmax, int H, int W);
/* Optional: build the per-tile seed lists of isg_assign_dense ahead of time (they depend on the seeds only, not on
 * kp / ae / thr_key), e.g. on the box branch while the top-k threshold is still being computed on another stream.
 * isg_assign_dense is then called with lists_prebuilt = 1 for the same seeds / workspace. */
int isg_build_tile_lists(const uint32_t* seeds, const int32_t* n_seeds, int B, int Nmax, int H, int W,
                         void* workspace, size_t workspace_bytes, isg_stream_t stream);
int isg_assign_dense(const float* kp, int64_t kp_img_stride,
                     const float* ae, int64_t ae_img_stride, int64_t ae_plane_stride,
                     const uint32_t* thr_key,
                     const uint32_t* seeds, const float* ghost, const int32_t* n_seeds, int B, int Nmax,
                     int H, int W, const float* ys, const float* xs,
                     int32_t* label_map, float* score_map, uint32_t* keepbits, int32_t* stats,
                     void* workspace, size_t workspace_bytes, int lists_prebuilt, isg_stream_t stream);
/* Split form of the dense step (kp is read from HBM once per step instead of twice):
 *   isg_topk_keep - isg_topk_threshold that also writes the keep bits (select_points, utils/decode.py:71-85: selected AND
 *     3x3 maximum of the thresholded map): the kernel that finds the exact threshold evaluates the peak test at the ~k
 *     selected pixels of its candidate list right away, instead of a second pass over the map (it walks the whole image by
 *     itself when the list is not complete).  keepbits [B,H,ceil(W/32)] is zeroed by the call unless keepbits_zeroed != 0
 *     (the caller zeroed it on this stream).  Same workspace as isg_topk_threshold.
 *   isg_assign_labels - isg_assign_dense without kp / thr_key / keepbits / stats: reads the 4 ae planes once, writes
 *     label_map (and score_map).  20 B/pixel.  Same workspace, tile lists and lists_prebuilt meaning as isg_assign_dense.
 *     Returns ISG_EUNSUPPORTED for layouts the tensor-map kernel cannot take (W % 4 != 0, unaligned planes): call
 *     isg_assign_dense instead.
 * Together they produce bit-identical label_map / keepbits to isg_assign_dense. */
int isg_topk_keep(const float* kp, int B, int H, int W, int64_t img_stride, int k, uint32_t* thr_key,
                  uint32_t* keepbits, int keepbits_zeroed, void* workspace, size_t workspace_bytes, isg_stream_t stream);
int isg_assign_labels(const float* ae, int64_t ae_img_stride, int64_t ae_plane_stride,
                      const uint32_t* seeds, const float* ghost, const int32_t* n_seeds, int B, int Nmax,
                      int H, int W, const float* ys, const float* xs, int32_t* label_map, float* score_map,
                      void* workspace, size_t workspace_bytes, int lists_prebuilt, isg_stream_t stream);
/* dense mode: labels / scores / ghost flags of the compacted keep pixels read back from the maps.
 * score_map nullable (then score is not written).  stats (nullable, pre-initialised by isg_stats_init): count /
 * bbox of the flagged pixels per instance - the same numbers isg_assign_dense accumulates when it is given a stats
 * pointer; pass it to exactly one of the two. */
int isg_gather_labels(const int32_t* label_map, const float* score_map, const int32_t* idx,
                      const int32_t* count, int cap, const float* ghost, int B, int Nmax, int H, int W,
                      int32_t* label, float* score, uint8_t* flag, int32_t* stats, isg_stream_t stream);
/* per-instance point sets (utils/decode.py:342-353): the flagged pixels of instance i, in row-major
 * order, as fp32 (x,y) pairs (detransform_pixel's flip, utils/tranform.py:157-159).
 * offsets: [B,Nmax+1] int32 exclusive prefix of the per-instance counts; points: [B,cap,2] fp32. */
int isg_group_points(const int32_t* idx, const int32_t* label, const uint8_t* flag, const int32_t* count,
                     int cap, const int32_t* n_seeds, int B, int Nmax,
                     int32_t* offsets, float* points, isg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Box head front-end.  Replaces BBoxTransform (utils/utils.py:318-346), ClipBoxes (:349-363) and
 * the score/threshold/class selection of decode_boxes (utils/decode.py:381-399).
 * anchors: [A,4] (y1,x1,y2,x2); regression: [B,A,4] (dy,dx,dh,dw); classification: [B,A,C].
 * Candidates (score > thr, fp32 compare) are appended in unspecified order; isg_box_nms orders
 * them by (score desc, anchor index asc).  cand_*: [B,cap,...]; cand_anchor: [B,cap] int32;
 * cand_count: [B] (true count, may exceed cap; only cap are stored); zeroed by the call.
 * ------------------------------------------------------------------------------------------ */
int isg_decode_boxes(const float* anchors, const float* regression, const float* classification,
                     int B, int A, int C, int H, int W, float thr, int cap,
                     float* cand_boxes, float* cand_scores, int32_t* cand_cls, int32_t* cand_anchor,
                     int32_t* cand_count, isg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K5 — greedy box NMS.  Replaces py_cpu_nms (utils/nms.py:11-39; ISG_NMS_PLUS1_LE) and
 * torchvision.ops.batched_nms as called at utils/decode.py:400 (ISG_NMS_TV_GT).
 * boxes [B,cap,4] (x1,y1,x2,y2), scores [B,cap], cls [B,cap] int32 (nullable = class agnostic),
 * tiebreak [B,cap] int32 (nullable: candidate index is used; on equal scores the LARGER tiebreak
 * value is visited first for PLUS1_LE and the SMALLER first for the TV_* conventions), count [B] device int32
 * (clamped to cap).  cap <= ISG_NMS_MAX_BOXES.  The three TV_* conventions differ only in the coordinates the IoU is
 * evaluated on (the shifted fp32 coordinates of the trick lose low-order bits, so a pair whose IoU is within ~1e-4 of
 * thr can resolve differently); boxes of different classes never suppress each other in any of them (with negative
 * coordinates torchvision's trick could let neighbouring classes overlap - not reproduced).
 * keep [B,cap] int32: candidate indices in pick order; n_keep [B].
 * ------------------------------------------------------------------------------------------ */
size_t isg_box_nms_workspace_bytes(int B, int cap);
int isg_box_nms(const float* boxes, const float* scores, const int32_t* cls, const int32_t* tiebreak,
                const int32_t* count, int B, int cap, double thr, int convention,
                int32_t* keep, int32_t* n_keep, void* ws, size_t ws_bytes, isg_stream_t stream);

/* BBoxTransform.forward for every anchor (utils/utils.py:318-346) -> boxes [B,A,4] (x1,y1,x2,y2); clip != 0
 * also applies ClipBoxes (utils/utils.py:349-363) for an H x W image. */
int isg_bbox_transform(const float* anchors, const float* regression, int B, int A, int clip, int H, int W,
                       float* boxes, isg_stream_t stream);
/* Anchors.forward (utils/utils.py:366-450): the multi-level anchor table [A,4] (y1,x1,y2,x2), bit-identical to the
 * reference's numpy-float64-then-astype construction.  strides [n_levels] HOST ints (the reference's 2**level);
 * half_sizes [n_levels][per_cell][2] HOST doubles = (anchor_size_x_2, anchor_size_y_2) of utils/utils.py:424-425 in
 * itertools.product(scales, ratios) order; n_levels <= 8, per_cell <= 16.  half_precision != 0 writes fp16 (the
 * reference's dtype == torch.float16 branch, :413-414).  isg_anchor_count is a host query (no GPU): A, or -1. */
int64_t isg_anchor_count(int H, int W, const int* strides, int n_levels, int per_cell);
int isg_generate_anchors(int H, int W, const int* strides, int n_levels, const double* half_sizes, int per_cell,
                         int half_precision, void* out /*[A,4]*/, isg_stream_t stream);
/* ClipBoxes.forward in place on n boxes (x1,y1,x2,y2) (utils/utils.py:357-361) */
int isg_clip_boxes(float* boxes, int64_t n, int H, int W, isg_stream_t stream);

/* kept candidates -> per-image detection tables in pick (= score descending) order, the dict that
 * decode_boxes returns (utils/decode.py:402-411) kept on the device: rois [B,Nmax,4], scores [B,Nmax],
 * cls [B,Nmax] int32, n_out [B] = min(n_keep, Nmax). */
int isg_gather_kept(const float* cand_boxes, const float* cand_scores, const int32_t* cand_cls,
                    const int32_t* keep, const int32_t* n_keep, int B, int cap, int Nmax,
                    float* rois, float* scores, int32_t* cls, int32_t* n_out, isg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K6 — bit-packed mask-IoU NMS.  IoU = (|A&B|+1)/(|A|B|+1) in fp64 (utils/image.py:188-191) inside
 * the greedy loop of utils/nms.py:23-37 (survivor iff IoU <= thr), class aware (cls nullable).
 * masks: [n,H,Wwords] uint32 (bit i of word w = pixel x = 32w+i; padding bits must be 0).
 * bboxes (nullable): [n,4] int32 (x0,y0,x1,y1) inclusive pixel bounds that contain every set bit of
 * the mask — used only to skip empty regions.
 * ------------------------------------------------------------------------------------------ */
size_t isg_mask_nms_workspace_bytes(int n);
int isg_mask_nms(const uint32_t* masks, int n, int H, int Wwords, const int32_t* bboxes,
                 const float* scores, const int32_t* cls, double thr,
                 int32_t* keep, int32_t* n_keep, void* ws, size_t ws_bytes, isg_stream_t stream);
/* dense masks [n,H,W] uint8 (non-zero = set; poly_to_mask output, utils/image.py:180-185) -> bit-packed
 * [n,H,ceil(W/32)] uint32 */
int isg_pack_masks(const uint8_t* dense, int n, int H, int W, uint32_t* bits, isg_stream_t stream);
/* pairwise mask statistics for n_pairs (a,b) index pairs: inter/union popcounts (int64 [n_pairs,2]).
 * Backs compute_iou_for_mask / is_cover (utils/image.py:188-191,205-207). */
int isg_mask_pair_counts(const uint32_t* masks, int n, int H, int Wwords, const int32_t* pairs, int n_pairs,
                         int64_t* inter_union, isg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K4 — seeded k-means with per-cluster allowed distance.  Replaces kmeans / pairwise_distance /
 * pairwise_cosine (utils/kmeans.py:16-130).  The whole Lloyd loop (assign + per-cluster sums, new
 * centres, `center_shift^2 < tol` test of utils/kmeans.py:90) is ONE cooperative launch; the call is
 * BLOCKING only at its end (it synchronises `stream` once to return the iteration count / status).
 * X [M,D] fp32, centers [N,D] fp32 in/out, allow [N] fp32, labels [M] int32 out (N = outlier; labels of the
 * LAST assignment, i.e. w.r.t. the pre-update centres, :93), iters_host: host int* (nullable).
 * max_iter <= 0: no bound (like the reference).  D <= 16; N*(12*D + 8) bytes of shared memory <= 200 KB.
 * The per-cluster mean is accumulated in fp64 in a fixed order (reproducible); the reference's fp32 mean
 * differs by ~1e-7 relative.
 * ------------------------------------------------------------------------------------------ */
size_t isg_kmeans_workspace_bytes(int M, int N, int D);
int isg_kmeans(const float* X, int M, int D, float* centers, const float* allow, int N,
               float tol, int metric, int max_iter, int32_t* labels, int* iters_host,
               void* ws, size_t ws_bytes, isg_stream_t stream);
/* pairwise_distance (:96-109) / pairwise_cosine (:112-130): out [M,N] fp32 */
int isg_pairwise(const float* X, int M, const float* Y, int N, int D, int metric, float* out,
                 isg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f1 — per-instance point sets + polygon extraction on the device (identity val-transform).  Replaces the per-instance
 * loop of group_kp (utils/decode.py:337-356) and aug_group (:167-204, with find_internal_point :51-68 and
 * cartesian2polar :88-113).  One CTA per (image, instance) collects, in row-major order, the keep pixels (keepbits,
 * label_map of isg_assign_dense) labelled with the instance that lie strictly inside its ghost bounds; instances with
 * at least obj_pixel_th points get their internal point, the polar-angle sort of their points (equal angles keep
 * row-major order; numpy's order among equal keys is unspecified) and the centre-inside test of the sorted polygon.
 *   rois [B,Nmax,4] fp32 in `layout` (ISG_BOX_XYXY / ISG_BOX_CYCXHW, as given to isg_build_seeds); ghost [B,Nmax,4] from isg_build_seeds; cap = capacity of poly_points per image
 *   poly_points [B,cap,2] fp32 (x,y): instance i of image b occupies [inst_start, inst_start+inst_count) of image b's
 *     block - angle-sorted when a polygon was computed, row-major otherwise; blocks are allocated in completion order
 *   inst_flags [B,Nmax] uint8: 1 = polygon valid (centre strictly inside), 0 = no polygon; instances with more than
 *     2048 points are finished by the same CTA in global memory (needs `workspace`); without a workspace
 *     they keep flag 2 and their raw row-major set, and the caller finishes them (aug_group on the host)
 *   inst_internal [B,Nmax,2] fp32 (nullable): the internal point used;  img_total [B] int32: points per image
 *   stats (nullable, pre-initialised by isg_stats_init): count / bbox per instance
 *   workspace (nullable): isg_instance_polygons_workspace_bytes(B, cap) bytes, 256-byte aligned
 *   totals_zeroed: 0 = img_total is zeroed by this call (one more small launch); 1 = the caller already zeroed it
 * ------------------------------------------------------------------------------------------ */
size_t isg_instance_polygons_workspace_bytes(int B, int cap);
int isg_instance_polygons(const uint32_t* keepbits, const int32_t* label_map, const float* rois, int layout, const float* ghost,
                          const int32_t* n_seeds, int B, int Nmax, int H, int W, int cap, int obj_pixel_th,
                          float* poly_points, int32_t* inst_start, int32_t* inst_count, uint8_t* inst_flags,
                          float* inst_internal, int32_t* img_total, int32_t* stats, void* workspace, size_t workspace_bytes,
                          int totals_zeroed, isg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f4 - the producer side: the 1x1 heads of EfficientDecoder at inference (models/efficient.py:508-510,536-541; heads
 * {"kp": 1, "ae": 4, "tan": 2}, :608).  Computes the five channels the decode reads (`tan` is ignored by decode_output,
 * utils/decode.py:447) in one pass over the decoder's last feature map x [B,Cin,H,W] (planar fp32, Cin <= 64, H*W % 4 == 0)
 * and writes them in the decode's own layout: kp [B,1,H,W], ae [B,4,H,W].
 * w_kp [1,Cin], b_kp [1], w_ae [4,Cin], b_ae [4]: HOST pointers (nn.Conv2d.weight / .bias of the kp and ae heads).
 * out[c] = bias[c] + sum_k w[c][k] * x[k] with fp32 FMAs in channel order (cuDNN / oneDNN order the sum differently:
 * equal to ~1e-6 relative, not bit for bit).
 * ------------------------------------------------------------------------------------------ */
int isg_decode_heads(const float* x, int B, int Cin, int H, int W, const float* w_kp, const float* b_kp,
                     const float* w_ae, const float* b_ae, float* kp, float* ae, isg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * One whole decode step = decode_output (utils/decode.py:444-461) for a batch already visible to the device,
 * enqueued by a single host call: isg_decode_boxes -> isg_box_nms -> isg_gather_build_seeds (-> isg_build_tile_lists)
 * on `main`; isg_topk_threshold (ISG_ASSIGN_SPARSE: + isg_keep_points + isg_compact_points) on `side`, forked from and
 * joined back into `main` with the two caller-owned events; then the assignment and (polygons != 0)
 * isg_instance_polygons on `main`.  Every pointer has the meaning documented at the entry point that consumes it.
 *   ISG_ASSIGN_DENSE : isg_assign_dense (label for every pixel; needs label_map, keepbits, dense_ws); with
 *                      split_keep != 0: isg_topk_keep on `side` instead of isg_topk_threshold + isg_assign_labels.
 *   ISG_ASSIGN_SPARSE: isg_assign_sparse + isg_scatter_labels (keep pixels only; needs idx, count, label too).  In this
 *                      mode `ae` and `regression` may be DEVICE ADDRESSES OF MAPPED HOST MEMORY (isg_host_device_pointer):
 *                      they are only read at the keep pixels / candidate anchors, so the planes never cross PCIe.
 * time_begin / time_end (nullable): events recorded on `main` around the assignment launches.
 * Nothing is synchronised; results are read by the caller after `main` has drained.  struct_bytes = sizeof(struct).
 * ------------------------------------------------------------------------------------------ */
#define ISG_ASSIGN_DENSE  0
#define ISG_ASSIGN_SPARSE 1
typedef struct isg_decode_step {
  int struct_bytes;
  int assign, polygons;
  int B, H, W, img_h, img_w, A, C, Nmax, cand_cap, cap, kp_th, obj_pixel_th;
  int nms_convention;          /* ISG_NMS_TV_GT / ISG_NMS_TV_TRICK / ISG_NMS_TV_BATCHED (utils/decode.py:400) */
  float cls_th, ghost_k, scale;
  double iou_th;
  /* model outputs */
  const float* kp; int64_t kp_img_stride;
  const float* ae; int64_t ae_img_stride, ae_plane_stride;
  const float* anchors; const float* regression; const float* classification;
  const float* ys; const float* xs;
  /* box head */
  float* cand_boxes; float* cand_scores; int32_t* cand_cls; int32_t* cand_anchor; int32_t* cand_count;
  int32_t* keep; int32_t* n_keep; void* nms_ws; size_t nms_ws_bytes;
  float* rois; float* scores; int32_t* cls; int32_t* n_seeds;
  /* selection + assignment */
  uint32_t* thr_key; void* topk_ws; size_t topk_ws_bytes;
  uint32_t* seeds; float* ghost; int32_t* stats;
  uint32_t* keepbits; int32_t* label_map; void* dense_ws; size_t dense_ws_bytes;
  int32_t* idx; int32_t* count; int32_t* label;
  /* per-instance polygons */
  float* poly_points; int32_t* inst_start; int32_t* inst_count; uint8_t* inst_flags; float* inst_internal;
  int32_t* img_total; void* poly_ws; size_t poly_ws_bytes;
  /* streams (cudaStream_t) and events (cudaEvent_t) */
  isg_stream_t main; isg_stream_t side; void* fork_event; void* join_event; void* time_begin; void* time_end;
  int split_keep;              /* ISG_ASSIGN_DENSE: keep bits from the top-k candidates, labels-only dense kernel */
} isg_decode_step_t;
int isg_decode_step(const isg_decode_step_t* step);
size_t isg_decode_step_bytes(void);   /* host query: sizeof(isg_decode_step_t) as compiled into the library */

/* ------------------------------------------------------------------------------------------
 * f2 - polygon rasteriser on the device.  Replaces poly_to_mask (utils/image.py:180-185 =
 * cv2.fillPoly(zeros(img_size, int32), [poly.astype(int32)], 1)) as called per detection by the results writer
 * (utils/eval_util.py:116), for n polygons per call, with bit-packed output.
 *   points [*,2] fp32 (x,y), 8-byte aligned; polygon i = points[poly_start[i] .. poly_start[i] + poly_count[i])
 *     (values are truncated to int32 like astype; every vertex must lie inside the H x W frame)
 *   full_frame = 1: polygon i fills words[i*H*Wwords ..) as an [H, Wwords = ceil(W/32)] frame - the layout of
 *     isg_mask_nms / isg_mask_pair_counts; full_frame = 0: polygon i gets rows x words_per_row words covering its
 *     bounding box (word aligned in x), allocated from `words` in completion order
 *   desc [n, ISG_FILL_DESC_WORDS] int32: {status, x0, y0, rows, words_per_row, offset_lo, offset_hi, n_vertices};
 *     bit k of word w of row r = pixel (x0 + 32 w + k, y0 + r)
 *   total: device scalar, zeroed by the call; words requested so far in compact mode (may exceed cap_words: the
 *     polygons that did not fit have status ISG_FILL_OVERFLOW)
 * ------------------------------------------------------------------------------------------ */
#define ISG_FILL_DESC_WORDS 8
#define ISG_FILL_OK       0
#define ISG_FILL_EMPTY    1   /* no vertices: nothing written */
#define ISG_FILL_OUTSIDE  2   /* a vertex lies outside the frame (OpenCV's edge clipping is not restated): nothing written */
#define ISG_FILL_OVERFLOW 3   /* `words` too small: nothing written for this polygon */
int isg_fill_polygons(const float* points, const int32_t* poly_start, const int32_t* poly_count, int n, int H, int W,
                      int full_frame, uint32_t* words, size_t cap_words, int32_t* desc, unsigned long long* total,
                      isg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * HOST helpers of the polygon stage (aug_group / find_internal_point, utils/decode.py:51-68,167-204).
 * Host pointers, no device work: the polygon stage is host glue this round (SURVEY.md §8 f1 is next).
 * They restate cv2.pointPolygonTest(contour fp32 [K,2], pt, measureDist=False) so that whole images are
 * processed per call instead of one cv2 call per test.
 * ------------------------------------------------------------------------------------------ */
/* +1 inside, 0 on the polyline, -1 outside */
int isg_host_point_in_polygon(const float* pts, int K, float px, float py);
/* find_internal_point for n instances; points [*,2] (x,y) grouped by instance, offsets [n+1], centers [n,2] (x,y),
 * internal [n,2] out.  Instances with fewer than min_pts points keep their centre. */
int isg_host_internal_points(const float* points, const int32_t* offsets, int n, const float* centers,
                             int min_pts, float* internal);
/* inside[i] = pointPolygonTest(polygon i, centers[i]) > 0 for n polygons stored back to back */
int isg_host_centres_inside(const float* points, const int32_t* offsets, int n, const float* centers,
                            uint8_t* inside);

#ifdef __cplusplus
}
#endif
#endif /* ISG_H_ */
