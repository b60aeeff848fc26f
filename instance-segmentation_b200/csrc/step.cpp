// isg_decode_step — one whole decode step (utils/decode.py:444-461: decode_boxes + per-image decode_single) enqueued by
// ONE host call: box head -> class-aware NMS -> detection tables + seeds -> tile lists on `main`, the top-k threshold
// on `side` (it does not depend on the boxes; split_keep: together with the keep bits, from its candidate list), then the
// assignment and the per-instance polygon stage on `main`.
// Pure host code over the entry points of include/isg.h: what engine.DecodePipeline did with a dozen Python calls per
// step, which bounded the step rate once neighbouring steps were overlapped.
#include <cuda_runtime.h>
#include "../../include/isg.h"

#define STEP_TRY(call)                 \
  do {                                 \
    const int rc__ = (call);           \
    if (rc__ != ISG_OK) return rc__;   \
  } while (0)
#define STEP_CUDA(call)                            \
  do {                                             \
    const cudaError_t e__ = (call);                \
    if (e__ != cudaSuccess) return (int)e__;       \
  } while (0)

extern "C" size_t isg_decode_step_bytes(void) { return sizeof(isg_decode_step_t); }

extern "C" int isg_decode_step(const isg_decode_step_t* s) {
  if (!s || s->struct_bytes != (int)sizeof(isg_decode_step_t)) return ISG_EINVAL;
  // `main` may be the default stream (a null handle); `side` must be a different stream
  if (!s->side || !s->fork_event || !s->join_event || s->main == s->side) return ISG_EINVAL;
  if (s->assign != ISG_ASSIGN_DENSE && s->assign != ISG_ASSIGN_SPARSE) return ISG_EINVAL;
  if (s->nms_convention != ISG_NMS_TV_GT && s->nms_convention != ISG_NMS_TV_TRICK && s->nms_convention != ISG_NMS_TV_BATCHED)
    return ISG_EINVAL;
  cudaStream_t main = (cudaStream_t)s->main, side = (cudaStream_t)s->side;
  cudaEvent_t fork = (cudaEvent_t)s->fork_event, join = (cudaEvent_t)s->join_event;
  const int B = s->B, H = s->H, W = s->W, N = s->Nmax;

  // top-k threshold on the side stream, behind everything enqueued on `main` so far
  STEP_CUDA(cudaEventRecord(fork, main));
  STEP_CUDA(cudaStreamWaitEvent(side, fork, 0));
  const bool split = s->assign == ISG_ASSIGN_DENSE && s->split_keep != 0;
  if (split) {    // threshold + keep bits from its candidate list in one go (the bit plane is zeroed first)
    STEP_CUDA(cudaMemsetAsync(s->keepbits, 0, (size_t)B * H * ((W + 31) / 32) * sizeof(uint32_t), side));
    STEP_TRY(isg_topk_keep(s->kp, B, H, W, s->kp_img_stride, s->kp_th, s->thr_key, s->keepbits, 1, s->topk_ws, s->topk_ws_bytes, side));
  } else {
    STEP_TRY(isg_topk_threshold(s->kp, B, H, W, s->kp_img_stride, s->kp_th, s->thr_key, s->topk_ws, s->topk_ws_bytes, side));
  }
  if (s->assign == ISG_ASSIGN_SPARSE) {   // keep bits + compaction only need kp and the threshold: stay on the side stream
    STEP_TRY(isg_keep_points(s->kp, B, H, W, s->kp_img_stride, s->thr_key, s->keepbits, nullptr, side));
    STEP_TRY(isg_compact_points(s->keepbits, B, H, W, s->cap, s->idx, s->count, side));
  }
  STEP_CUDA(cudaEventRecord(join, side));

  // box branch
  STEP_TRY(isg_decode_boxes(s->anchors, s->regression, s->classification, B, s->A, s->C, s->img_h, s->img_w, s->cls_th,
                            s->cand_cap, s->cand_boxes, s->cand_scores, s->cand_cls, s->cand_anchor, s->cand_count, main));
  STEP_TRY(isg_box_nms(s->cand_boxes, s->cand_scores, s->cand_cls, s->cand_anchor, s->cand_count, B, s->cand_cap, s->iou_th,
                       s->nms_convention, s->keep, s->n_keep, s->nms_ws, s->nms_ws_bytes, main));
  STEP_TRY(isg_gather_build_seeds(s->cand_boxes, s->cand_scores, s->cand_cls, s->keep, s->n_keep, B, s->cand_cap, N, s->ys,
                                  s->xs, H, W, s->ghost_k, s->scale, s->rois, s->scores, s->cls, s->n_seeds, s->seeds,
                                  s->ghost, s->stats, s->img_total, main));
  if (s->assign == ISG_ASSIGN_DENSE)
    STEP_TRY(isg_build_tile_lists(s->seeds, s->n_seeds, B, N, H, W, s->dense_ws, s->dense_ws_bytes, main));
  STEP_CUDA(cudaStreamWaitEvent(main, join, 0));

  if (s->time_begin) STEP_CUDA(cudaEventRecord((cudaEvent_t)s->time_begin, main));
  if (s->assign == ISG_ASSIGN_DENSE) {
    int rc = ISG_EUNSUPPORTED;
    if (split)
      rc = isg_assign_labels(s->ae, s->ae_img_stride, s->ae_plane_stride, s->seeds, s->ghost, s->n_seeds, B, N, H, W, s->ys,
                             s->xs, s->label_map, nullptr, s->dense_ws, s->dense_ws_bytes, 1, main);
    if (rc == ISG_EUNSUPPORTED)     // not split, or a layout only the fused form takes (it rewrites the same keep bits)
      rc = isg_assign_dense(s->kp, s->kp_img_stride, s->ae, s->ae_img_stride, s->ae_plane_stride, s->thr_key, s->seeds,
                            s->ghost, s->n_seeds, B, N, H, W, s->ys, s->xs, s->label_map, nullptr, s->keepbits, nullptr,
                            s->dense_ws, s->dense_ws_bytes, 1, main);
    STEP_TRY(rc);
  } else {
    // only the keep pixels: `ae` may live in mapped host memory (16 B per keep pixel cross PCIe instead of the planes)
    STEP_TRY(isg_assign_sparse(s->ae, s->ae_img_stride, s->ae_plane_stride, s->idx, s->count, s->cap, s->seeds, s->ghost,
                               s->n_seeds, B, N, H, W, s->ys, s->xs, s->label, nullptr, nullptr, nullptr, main));
    STEP_TRY(isg_scatter_labels(s->idx, s->count, s->cap, s->label, B, H, W, s->label_map, main));
  }
  if (s->time_end) STEP_CUDA(cudaEventRecord((cudaEvent_t)s->time_end, main));
  if (s->polygons)
    STEP_TRY(isg_instance_polygons(s->keepbits, s->label_map, s->rois, ISG_BOX_XYXY, s->ghost, s->n_seeds, B, N, H, W, s->cap,
                                   s->obj_pixel_th, s->poly_points, s->inst_start, s->inst_count, s->inst_flags,
                                   s->inst_internal, s->img_total, s->stats, s->poly_ws, s->poly_ws_bytes, 1, main));
  return ISG_OK;
}
