/* ut2.h — C ABI of libut2_sm100.so, the B200 (sm_100a) kernel library behind the Unbiased-Teacher-v2 hot path.
 *
 * The reference has no FFI of its own: every device op below is reached there through Python -> torch /
 * Detectron2 / torchvision / fvcore. Each entry point cites the reference call site it replaces
 * (paths under /root/reference/ubteacher unless marked [D2] = Detectron2 v0.6, [tv] = torchvision, [fvcore]).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; all pointers are DEVICE pointers unless a parameter says "host".
 *   - the caller owns every buffer (outputs and workspaces); the library never allocates device memory.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*) and never synchronises the host.
 *   - return value: 0 = ok, < 0 = argument / shape error, > 0 = cudaError_t. ut2_last_error_string() (thread-local)
 *     describes the last failure. No exceptions cross the boundary.
 *   - activations are NHWC bf16 (== row-major [N*H*W, C]); conv weights are [Cout, R, S, Cin] bf16; accumulation fp32.
 *   - variable-length results are fixed-capacity buffers + a device-side count.
 *   - "level-major" per-location tensors: position p = level_off[l]*N + img*H_l*W_l + h*W_l + w, i.e. the per-level
 *     NHWC head outputs laid back to back (the order fcos_outputs.py:257-296 builds with permute+reshape+cat).
 */
#ifndef UT2_H
#define UT2_H
#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- library */
int ut2_version(void);
const char* ut2_last_error_string(void);
int ut2_device_sm_count(void);

/* ---------------------------------------------------------------- tensor-core convolutions (tcgen05 + TMA im2col)
 * Replaces torch.nn.functional.conv2d -> cuDNN for: [D2] ResNet bottlenecks / FPN (modeling/backbone/fpn.py:59-78),
 * LastLevelP6P7 (backbone/fpn.py:11-29), FCOS towers and prediction convs (modeling/fcos/fcos.py:248-304,338-376).
 * y[n,p,q,co] = act( scale[co] * sum_{r,s,ci} x[n, p*stride-pad+r, q*stride-pad+s, ci] * w[co,r,s,ci] + shift[co]
 *                    (+ residual[n,p,q,co] | residual[n,p/2,q/2,co] when res_up2) ) , then zeroed where relu_mask <= 0.
 * Cin % 8 == 0; Cout % 16 == 0 (<= 256, or a multiple of 128). scale/shift/residual/relu_mask may be NULL.
 * The data-gradient is the same entry point applied to dY with the flipped/transposed filter (see ut2_pack_*). */
int ut2_conv2d_nhwc_bf16_fwd(const void* x, int N, int H, int W, int Cin, const void* w, int Cout, int R, int S,
                             int stride, int pad, const float* scale, const float* shift, const void* residual,
                             int res_up2, const void* relu_mask, int relu, void* y, void* stream);
/* 3x3 / stride 1 / pad 1 launches with 64, 80 or 128 output channels, no scale vector and no residual (the res2 / res3 conv2 of the
 * bottlenecks, their data-gradients, the FCOS predictors through ut2_conv2d_levels_bf16_fwd) run on 2-D output patches with tiled-mode TMA (csrc/conv3x3_halo.cu: each input slice is
 * fetched 3x instead of 9x) when they have at least one patch per SM; same results up to fp32 summation order.
 * UT2_HALO3=0 disables. Number of such launches so far (tests check that the path is taken): */
long long ut2_conv3x3_halo_launches(void);
/* dW[co,r,s,ci] (fp32, accumulated with atomics) += scale[co] * sum_{n,p,q} dy[n,p,q,co] * x[n,p*stride-pad+r,...,ci].
 * Replaces the cuDNN wgrad autograd runs for the same layers (engine/trainer.py:422-429 -> losses.backward()).
 * Rows >= cout_store (if > 0) are not written (zero-padded fused predictors). Cin % 64 == 0, Cout % 8 == 0. */
int ut2_conv2d_nhwc_bf16_wgrad(const void* x, int N, int H, int W, int Cin, const void* dy, int Cout, int R, int S,
                               int stride, int pad, const float* scale, float* dw, int cout_store, void* stream);
/* The same two operators over a level-major pyramid in ONE launch: x = levels [N, H_l, W_l, Cin] laid back to back
 * (rows sum_l N*H_l*W_l), y / dy likewise; hw = HOST array {H_0, W_0, H_1, W_1, ...}, <= 5 levels, stride 1. Used for the
 * FCOS towers and predictors, which apply the same weights to all five FPN levels (fcos/fcos.py:338-376). */
int ut2_conv2d_levels_bf16_fwd(const void* x, int num_levels, const int* hw, int N, int Cin, const void* w, int Cout, int R,
                               int S, int pad, const float* scale, const float* shift, const void* residual,
                               const void* relu_mask, int relu, void* y, void* stream);
int ut2_conv2d_levels_bf16_wgrad(const void* x, int num_levels, const int* hw, int N, int Cin, const void* dy, int Cout,
                                 int R, int S, int pad, const float* scale, float* dw, int cout_store, void* stream);
/* test hook: one im2col-mode TMA load dumped from shared memory (pins the descriptor semantics) */
int ut2_debug_im2col_probe(const void* x, int N, int H, int W, int C, int R, int S, int stride, int pad, int pixels,
                           int c, int w, int h, int n, int off_w, int off_h, void* out, void* stream);

/* ---------------------------------------------------------------- backbone helpers
 * stem: pixel normalisation (one_stage_detector.py:88-89,165-166) + ImageList zero padding (:90,:167) + [D2] BasicStem
 * conv7x7 s2 + FrozenBN + ReLU, from one uint8 CHW (BGR) image; wgt is fp32 [7][7][3][64]. */
int ut2_stem_conv_u8(const void* img_chw, int h, int w, const float* wgt_rsck, const float* scale, const float* shift,
                     float mean0, float mean1, float mean2, float std0, float std1, float std2, void* out, int P, int Q,
                     void* stream);
/* same contract on the tensor cores (implicit GEMM, K = 147 padded to 192, bf16 inputs, fp32 accumulate) */
int ut2_stem_conv_u8_tc(const void* img_chw, int h, int w, const float* wgt_rsck, const float* scale, const float* shift,
                        float mean0, float mean1, float mean2, float std0, float std1, float std2, void* out, int P, int Q,
                        void* stream);
/* the same over a batch in one launch: imgs / hs / ws are HOST arrays of N device pointers (uint8 CHW) and image sizes;
 * out is [N, P, Q, 64] bf16 (P, Q = padded size / 2: ImageList.from_tensors zero padding after normalisation). */
int ut2_stem_conv_u8_tc_batched(const void* const* imgs, const int* hs, const int* ws, int N, const float* wgt_rsck,
                                const float* scale, const float* shift, float m0, float m1, float m2, float s0, float s1,
                                float s2, void* out, int P, int Q, void* stream);
/* Fused stem + max-pool (uint8 CHW images -> [N, Hp/4, Wp/4, 64] bf16): normalise + space-to-depth scratch, then the 7x7 s2
 * convolution as a 4x4 s1 tcgen05 implicit GEMM with FrozenBN / ReLU / 3x3 s2 max-pool in the epilogue; the 64-channel stem
 * activation never reaches HBM. Replaces ut2_stem_conv_u8_tc_batched + ut2_maxpool3x3s2_nhwc ([D2] BasicStem). imgs / hs / ws:
 * HOST arrays; hw_dev: optional DEVICE int[N][2] with the same (h, w), read by the kernel instead of hs / ws — the sizes are then
 * data, and a captured CUDA graph of the step serves every batch of the same padded size; workspace:
 * ut2_stem_pool_workspace_bytes(N, Hp, Wp) device bytes. */
long long ut2_stem_pool_workspace_bytes(int N, int Hp, int Wp);
int ut2_stem_pool_u8_batched(const void* const* imgs, const int* hs, const int* ws, int N, const float* wgt_rsck,
                             const float* scale, const float* shift, float m0, float m1, float m2, float s0, float s1, float s2,
                             void* workspace, long long workspace_bytes, void* out, int Hp, int Wp, const int* hw_dev, void* stream);
int ut2_maxpool3x3s2_nhwc(const void* x, void* y, int N, int H, int W, int C, void* stream);          /* [D2] BasicStem max_pool2d */
int ut2_upsample2x_add_nhwc(const void* lat, const void* top, void* out, int N, int H, int W, int C, void* stream); /* [D2] FPN top-down */
int ut2_downsample2x_sum_nhwc(const void* g, const void* addend, void* gtop, int N, int Ht, int Wt, int C, void* stream);
int ut2_relu_bwd_bf16(const void* dy, const void* dy2, const void* y, void* g, long long n, void* stream);
int ut2_add_bf16(const void* a, const void* b, void* out, long long n, void* stream);
int ut2_zero_stuff_s2_nhwc(const void* in, void* out, int N, int P, int Q, int H, int W, int C, int oh, int ow, void* stream);
/* SM budget of the persistent conv kernels: lowered by the trainer while NCCL all-reduces of finished gradient segments run
 * on a side stream (DDP's bucketed all-reduce inside backward, reference engine/trainer.py:60-63,428). 0 = whole device. */
int ut2_set_sm_limit(int n);
int ut2_colsum_bf16(const void* g, float* db, int M, int C, void* stream);                             /* conv bias gradients */
/* the same for up to 16 (gradient matrix [M_i, C_i], bias gradient) pairs in one launch; HOST arrays of device pointers */
int ut2_colsum_bf16_batched(const void* const* gs, float* const* dbs, const int* Ms, const int* Cs, int n, void* stream);
int ut2_frozen_bn_fold(const float* w, const float* b, const float* mean, const float* var, float eps, float* scale,
                       float* shift, int C, void* stream);                                             /* [D2] FrozenBatchNorm2d */
int ut2_cast_f32_bf16(const float* x, void* y, long long n, void* stream);

/* GroupNorm(32, 256) + ReLU of the FCOS towers (fcos/fcos.py:263-264,283). stats/ws: double[N*32*2]. */
int ut2_groupnorm_relu_fwd(const void* x, const float* gamma, const float* beta, float eps, void* y, double* stats,
                           int N, int HW, int C, int G, int relu, void* stream);
int ut2_groupnorm_relu_bwd(const void* dy, const void* x, const double* stats, const float* gamma, const float* beta,
                           float eps, void* dx, float* dgamma, float* dbeta, float* dbias_prev /* += colsum(dx), optional */,
                           double* ws, int N, int HW, int C, int G, int relu, void* stream);

/* level-major variants (hws = HOST array of H_l*W_l; statistics per (level, image, group); stats / ws double[levels*N*32*2]) */
int ut2_groupnorm_relu_levels_fwd(const void* x, const float* gamma, const float* beta, float eps, void* y, double* stats,
                                  int num_levels, const int* hws, int N, int C, int G, int relu, void* stream);
int ut2_groupnorm_relu_levels_bwd(const void* dy, const void* x, const double* stats, const float* gamma, const float* beta,
                                  float eps, void* dx, float* dgamma, float* dbeta, float* dbias_prev, double* ws,
                                  int num_levels, const int* hws, int N, int C, int G, int relu, void* stream);

/* Stand-alone loss operators behind ubteacher.layers.{IOULoss, NLLoss, KLLoss} (ubteacher/layers/iou_loss.py:23-76,
 * kl_loss.py:17-105); rows [P, 4] f32, value in loss[0], input gradients (d loss / d input) written when the pointers are given.
 * The training step computes the same terms inside ut2_fcos_loss_fwd / _bwd. */
int ut2_iou_loss(const float* pred, const float* target, const float* weight /* [P] or NULL */, int P, int type /* 0 iou, 1 linear_iou, 2 giou */,
                 double* acc /* [1] scratch */, float* loss, float* dpred /* [P,4] or NULL */, void* stream);
int ut2_nl_loss(const float* mean, const float* std, const float* target, const float* iou_weight, int P, double* acc, float* loss,
                float* dmean, float* dstd, void* stream);
int ut2_kl_loss(const float* input, const float* std, const float* target, const float* weight, int P, float beta,
                int method /* 0 weight_ctr_sum, 1 weight_ctr_mean, 2 sum, 3 mean */, float loss_denorm, double* acc, float* loss,
                float* dinput, float* dstd, void* stream);
int ut2_scale_f32(const float* x, const float* s /* device scalar */, float* y, long long n, void* stream);

/* ---------------------------------------------------------------- FCOS targets and losses
 * ut2_fcos_assign_targets: FCOSOutputs._get_ground_truth + compute_targets_for_locations
 * (fcos/fcos_outputs.py:649-698, :772-906). hw/strides/ranges are HOST arrays
 * ([H,W] x levels, stride x levels, [lo,hi] x levels). boxes [N,G,4], classes [N,G] (int64), counts [N] (int32),
 * bvar [N,G,4] or NULL (teacher reg_pred_std). center_radius > 0: MODEL.FCOS.CENTER_SAMPLE with POS_RADIUS = center_radius
 * (get_sample_region, :700-770); 0: a location is positive anywhere inside the box (the shipped recipes). ignore_near != 0:
 * keep[] drops the locations inside a box but outside every sample region (:841-848). norm[2] receives {num_pos, sum of
 * centerness targets}. */
int ut2_fcos_assign_targets(int num_levels, const int* hw, const int* strides, const float* ranges, int N, int G,
                            const float* boxes, const long long* classes, const int* counts, const float* bvar,
                            int num_classes, float center_radius, int ignore_near, long long* labels, long long* tinds,
                            float* reg_t, float* bv_out, unsigned char* keep, float* norm, void* stream);
/* mode 0: FCOSOutputs.fcos_losses (fcos_outputs.py:307-444); mode 1 / 2: fcos_pseudo_losses on the classification /
 * regression pseudo-label set (:492-631). Includes Integral (:44-77), centerness / IoU targets (:80-129), IOULoss giou
 * (layers/iou_loss.py:23-76), NLLoss (layers/kl_loss.py:75-105), [fvcore] sigmoid_focal_loss_jit.
 * cls_out / box_out are [P, ld] bf16 (box_out: 68 distribution logits | 4 std | 1 centerness | pad); scales[levels] is
 * the learnable Scale (fcos/fcos.py:22-29,367). norm is the (all-reduced) output of ut2_fcos_assign_targets, world the
 * number of ranks. kl_mode: MODEL.FCOS.KL_LOSS_TYPE / LOC_FUN_ALL of the supervised branch (:377-416) — 0 "nlloss" (the shipped
 * recipes), 1..4 "klloss" reduced by "mean" / "sum" / "weight_ctr_sum" / "weight_ctr_mean" (layers/kl_loss.py:17-66).
 * acc: double[8] scratch kept for backward; losses: float[4] = {cls, loc, ctr, teacher_better_student}. */
int ut2_fcos_loss_fwd(int num_levels, const int* hw, const int* strides, int N, const void* cls_out, const void* box_out,
                      int ld, const float* scales, const long long* labels, const unsigned char* keep, const float* reg_t,
                      const float* bvar, int num_classes, int mode, float alpha, float gamma, float kl_w, float ts_better,
                      float ts_cert, const float* norm, float world, int kl_mode, double* acc, float* losses, void* stream);
int ut2_fcos_loss_bwd(int num_levels, const int* hw, const int* strides, int N, const void* cls_out, const void* box_out,
                      int ld, const float* scales, const long long* labels, const unsigned char* keep, const float* reg_t,
                      const float* bvar, int num_classes, int mode, float alpha, float gamma, float kl_w, float ts_better,
                      float ts_cert, const float* norm, float world, int kl_mode, const double* acc, const float* gout, void* dcls,
                      void* dbox, float* dscales, int accumulate, void* stream);

/* ---------------------------------------------------------------- proposals, NMS, pseudo labels
 * FCOSOutputs.predict_proposals + forward_for_single_feature_map + select_over_all_levels
 * (fcos_outputs.py:1046-1320) with ml_nms -> [D2] batched_nms -> [tv] nms (layers/ml_nms.py:8-31), bit-exact
 * coordinate-trick IoU. method 0 "cls", 1 "cls_n_ctr", 2 "cls_n_loc". Outputs are [N, out_cap(, k)] + out_cnt[N]. */
long long ut2_fcos_predict_workspace_bytes(int num_levels, int N, long long L, int C, int K);
int ut2_fcos_predict_proposals(int num_levels, const int* hw, const int* strides, int N, int C, const void* cls_out,
                               const void* box_out, int ld, const float* scales, int method, float pre_thr, int pre_topk,
                               float nms_thr, int post_topk, int out_cap, void* workspace, long long workspace_bytes,
                               float* out_box, float* out_score, long long* out_cls, float* out_ctr, float* out_conf,
                               float* out_std, float* out_loc, long long* out_lvl, int* out_cnt, void* stream);
/* PseudoGenerator.threshold_bbox (mode 0: score > thr0) / threshold_cls_ctr_bbox (mode 1: cls_confid > thr0 and
 * centerness > thr1) — modeling/pseudo_generator.py:62-131; order-preserving compaction. */
int ut2_threshold_scatter(int N, int cap, int mode, float thr0, float thr1, const int* in_cnt, const float* box,
                          const float* score, const long long* cls, const float* ctr, const float* conf,
                          const float* stdv, int* out_cnt, float* obox, float* oscore, long long* ocls, float* octr,
                          float* oconf, float* ostd, void* stream);

/* ---------------------------------------------------------------- optimiser / EMA over flat arenas
 * _update_teacher_model (engine/trainer.py:468-486): teacher = student*(1-keep) + teacher*keep, bit-exact with torch
 * (both products rounded to fp32, one rounded add). */
int ut2_ema_update(const float* student, float* teacher, long long n, double keep_rate, void* stream);
/* torch.optim.SGD as built by [D2] build_optimizer (engine/trainer.py:422-429): g' = grad_scale*g + wd*p;
 * buf = first ? g' : mom*buf + g'; p -= lr*buf; optionally clears g. */
int ut2_sgd_step(float* p, float* g, float* buf, long long n, float lr, const float* lr_dev /* overrides lr if set */,
                 float momentum, float weight_decay, int first_step, int zero_grad, float grad_scale, void* stream);
/* fp32 master weights (channels-last) -> bf16 forward operand [Cout,R,S,Cin] and dgrad operand [Cin,R,S,CoutT]
 * (taps flipped). The batched form takes a device table of 64-byte records
 * {int64 src, wf, wt, begin, scale; int32 Cout, Cin, R, S, CoutT, n_off}; begin = prefix sum of the records' tile counts
 * R*S*ceil(Cout/64)*ceil(Cin/32), total_tiles its end; scale >= 0 folds scales[scale + co] (the FrozenBN scale) into
 * the packed weights; write_dgrad = 0 leaves the dgrad operands untouched (inference-only replica: the EMA teacher). */
int ut2_pack_conv_weight(const float* w, void* wf, void* wt, int Cout, int Cin, int R, int S, int CoutT, void* stream);
int ut2_pack_conv_weights_batched(const void* descs, int num, long long total_tiles, const float* arena,
                                  const float* scales, void* packed, int write_dgrad, void* stream);

/* ================================================================ Faster R-CNN half
 * [D2] = Detectron2 v0.6 (not on disk; behaviour restated in SURVEY.md appendix B). */

/* ---------------------------------------------------------------- generic batched NMS
 * [D2] batched_nms -> [tv] batched_nms/nms as called by [D2] find_top_rpn_proposals (modeling/proposal_generator/
 * rpn.py:72-74) and [D2] fast_rcnn_inference (modeling/roi_heads/fast_rcnn.py:1112-1119). boxes [N,M,4], scores [N,M]
 * (non-finite = dropped), tie [N,M] tie-break key (NULL: slot index), cls [N,M], cnt [N] used slots. Visits candidates
 * by (score desc, tie asc); coordinate trick iff 4*n_valid <= trick_limit else per-class on raw boxes (torchvision's
 * switch). keep_idx [N,max_keep] = slot indices of the first max_keep survivors, keep_cnt [N]. M <= 16384. */
long long ut2_nms_workspace_bytes(int N, int M);
int ut2_nms_batched(int N, int M, const float* boxes, const float* scores, const int* tie, const int* cls, const int* cnt,
                    float thr, int trick_limit, int max_keep, void* workspace, long long workspace_bytes, int* keep_idx,
                    int* keep_cnt, void* stream);
/* the same for candidate lists laid out class by class — the RPN's level-by-level list ([D2] find_top_rpn_proposals, idxs =
 * FPN level): seg_off = HOST array of S + 1 slot offsets inside every image's M slots (S <= 8, segments <= 2048 slots, all
 * slots candidates; non-finite scores dropped). Each (image, segment) is sorted / masked / scanned on its own and the
 * survivors merged by (score desc, tie asc): identical keep list to ut2_nms_batched with cls = segment index. */
long long ut2_nms_segmented_workspace_bytes(int N, int S, int Mseg, int max_keep);
int ut2_nms_segmented(int N, int M, int S, const int* seg_off, const float* boxes, const float* scores, const int* tie, float thr,
                      int trick_limit, int max_keep, void* workspace, long long workspace_bytes, int* keep_idx, int* keep_cnt,
                      void* stream);
/* dst[img,k,:] = src[img, idx[img,k], :] for k < cnt[img], else 0; rows of W elements of elem_bytes (4 | 8) */
int ut2_gather_rows(int N, int M, int K, int W, int elem_bytes, const void* src, const int* idx, const int* cnt, void* dst,
                    void* stream);

/* ---------------------------------------------------------------- RPN (modeling/proposal_generator/rpn.py)
 * rpn_out: fused predictor output, level-major [N*sum(H_l*W_l), 16] bf16 (0..2 objectness, 3+4a+k deltas, 15 pad);
 * hw / strides / cell (cell anchors [levels][3][4], [D2] DefaultAnchorGenerator) are HOST arrays.
 * ut2_rpn_label_anchors: label_and_sample_anchors[_pseudo] (rpn.py:78-150): [D2] pairwise_iou + Matcher(lo, hi,
 * allow_low_quality) + subsample_labels; the randperm draw is a per-anchor uint32 key (keys [N,A] or hashed from seed;
 * seed_dev: optional DEVICE word mixed into the seed so that a captured CUDA graph draws afresh on every replay):
 * the smallest (key, index) win. labels int8 [N,A] in {-1,0,1}; matched int32 [N,A] = argmax ground truth. */
long long ut2_rpn_label_workspace_bytes(int N, long long A, int G);
int ut2_rpn_label_anchors(int num_levels, const int* hw, const int* strides, const float* cell, int N, int G,
                          const float* gt_boxes, const int* gt_cnt, const unsigned int* keys, unsigned int seed,
                          const unsigned int* seed_dev, int batch_per_image, float pos_fraction, float lo_thr, float hi_thr, void* workspace,
                          long long workspace_bytes, signed char* labels, int* matched, void* stream);
/* PseudoLabRPN.losses (rpn.py:153-225): BCE-with-logits over sampled anchors (x matched teacher score when gt_scores
 * is given) and L1 on the positives' deltas ([D2] _dense_box_regression_loss, Box2BoxTransform weights 1), both
 * / (batch_per_image * N). acc double[2]; losses float[2] = {loss_rpn_cls, loss_rpn_loc}; gout float[2]. */
int ut2_rpn_loss_fwd(int num_levels, const int* hw, const int* strides, const float* cell, int N, int G, const void* rpn_out,
                     const signed char* labels, const int* matched, const float* gt_boxes, const float* gt_scores,
                     const int* gt_cnt, int batch_per_image, double* acc, float* losses, void* stream);
int ut2_rpn_loss_bwd(int num_levels, const int* hw, const int* strides, const float* cell, int N, int G, const void* rpn_out,
                     const signed char* labels, const int* matched, const float* gt_boxes, const float* gt_scores,
                     const int* gt_cnt, int batch_per_image, const float* gout, void* drpn, void* stream);
/* first half of [D2] find_top_rpn_proposals (rpn.py:72-74): per (image, level) top pre_topk logits (ties: lower anchor
 * index), Box2BoxTransform.apply_deltas, clip to image_hw [N,2] (device), invalid -> score -inf. Feed to ut2_nms_batched. */
int ut2_rpn_select_decode(int num_levels, const int* hw, const int* strides, const float* cell, int N, const void* rpn_out,
                          const float* image_hw, int pre_topk, float scale_clamp, int Mcap, float* cand_box,
                          float* cand_score, int* cand_canon, int* cand_lvl, int* cand_cnt, void* stream);

/* ---------------------------------------------------------------- ROI heads (modeling/roi_heads/roi_heads.py, fast_rcnn.py)
 * ut2_roi_sample: label_and_sample_proposals[_pseudo] (roi_heads.py:138-270) — append GT, Matcher(iou_thr), sample Rcap
 * ROIs with <= pos_fraction foreground by (key, index); copies gt class / box and, for pseudo labels, the teacher score
 * (gt_confid) and box std (gt_loc_std). Outputs [N,Rcap,...] fg first; rows >= roi_cnt[img] have class -1. */
int ut2_roi_sample(int N, int Pcap, int G, int Rcap, const float* prop_boxes, const int* prop_cnt, const float* gt_boxes,
                   const long long* gt_classes, const int* gt_cnt, const float* gt_scores, const float* gt_std,
                   const unsigned int* keys, int key_ld, unsigned int seed, const unsigned int* seed_dev, float pos_fraction,
                   float iou_thr, int num_classes, int append_gt, float* roi_box, long long* roi_cls, float* roi_gtbox, float* roi_conf, float* roi_std,
                   int* roi_src, int* roi_cnt, void* stream);
/* [D2] ROIPooler(7x7, ROIAlignV2, sampling_ratio 0) -> [tv] roi_align(aligned=True) (roi_heads.py:118). feats / dfeats:
 * HOST arrays of per-level device pointers (NHWC bf16 / fp32 accumulators); out / dout [N*Rcap, 7, 7, C] bf16. */
int ut2_roi_align_fwd(int num_levels, const void* const* feats, const int* hw, const float* scales, int N, int C, int Rcap,
                      const float* rois, const int* roi_cnt, void* out, void* stream);
int ut2_roi_align_bwd(int num_levels, float* const* dfeats, const int* hw, const float* scales, int N, int C, int Rcap,
                      const float* rois, const int* roi_cnt, const void* dout, void* stream);
/* FastRCNNFocaltLossBoundaryVarOutputLayers.losses (fast_rcnn.py:834-1084): FocalLoss(gamma)/R; mode 0 'nlloss' box loss
 * [L1 + nll_w * NLL * IoU(pred, gt)]/R, mode 1 'tsbetter' teacher-better-than-student masked L1 / R. pred [N*Rcap, 96]
 * bf16 (0..80 scores | 81..84 deltas | 85..88 std). acc double[2], losses float[2] = {loss_cls, loss_box_reg}. */
int ut2_fastrcnn_loss_fwd(int N, int Rcap, const void* pred, const float* rois, const long long* gt_cls, const float* gt_box,
                          const float* gt_std, const int* roi_cnt, int mode, float wx, float wy, float clamp, float gamma,
                          float nll_w, float ts_better, float t_cert, double* acc, float* losses, void* stream);
int ut2_fastrcnn_loss_bwd(int N, int Rcap, const void* pred, const float* rois, const long long* gt_cls, const float* gt_box,
                          const float* gt_std, const int* roi_cnt, int mode, float wx, float wy, float clamp, float gamma,
                          float nll_w, float ts_better, float t_cert, const float* gout, void* dpred, void* stream);
/* inference (fast_rcnn.py:1086-1125 -> [D2] fast_rcnn_inference): softmax, Box2BoxXYXYTransform.apply_deltas, clip,
 * p > score_thr -> candidates (then ut2_nms_batched by class) -> ut2_fastrcnn_gather (adds pred_boxes_std, :1122-1123). */
int ut2_fastrcnn_candidates(int N, int Rcap, const void* pred, const float* rois, const int* roi_cnt, const float* image_hw,
                            float wx, float wy, float clamp, float score_thr, int Ccap, float* cand_box, float* cand_score,
                            int* cand_cls, int* cand_canon, int* cand_cnt, int* overflow, void* stream);
int ut2_fastrcnn_gather(int N, int Ccap, int K, int Rcap, const int* keep_idx, const int* keep_cnt, const float* cand_box,
                        const float* cand_score, const int* cand_canon, const void* pred, float* out_box, float* out_score,
                        long long* out_cls, float* out_std, int* out_roi, int* out_cnt, void* stream);
/* Box2BoxXYXYTransform as a stand-alone operator (ubteacher/modeling/box_regression.py:36-75 get_deltas, :77-129
 * apply_deltas; the loss / inference kernels above apply the same arithmetic in registers). Boxes [n,4] xyxy f32;
 * deltas (dl, dr, dd, du) [n, 4*k]; get_deltas divides by size + 1, apply_deltas multiplies by size (reference quirk). */
int ut2_box2box_xyxy_get_deltas(const float* src_boxes, const float* target_boxes, int n, float wx, float wy, float* deltas,
                                void* stream);
int ut2_box2box_xyxy_apply_deltas(const float* deltas, const float* boxes, int n, int k, float wx, float wy, float clamp,
                                  float* pred_boxes, void* stream);
int ut2_add_f32_bf16(const float* a, const void* b /* bf16, optional */, void* out, long long n, void* stream);
int ut2_subsample2x_nhwc(const void* x, void* y, int N, int H, int W, int C, void* stream);   /* [D2] LastLevelMaxPool */

/* ================================================================ two-crop strong augmentation
 * ubteacher/data/detection_utils.py:8-46 (torchvision ColorJitter / RandomGrayscale / GaussianBlur / RandomErasing on PIL
 * images), restated bit for bit on planar uint8 [3,h,w] device images. All random draws arrive in `table`, a DEVICE
 * array of N 168-byte records (little endian, python struct "<6Q2i4i4f3i2Ii12iIi"):
 *   u64 src, dst, tmp (blur scratch), noise[3] (optional float32 [3,eh,ew] per erase; 0 = hashed N(0,1));
 *   i32 h, w; i32 order[4] (ColorJitter op per slot: 0 brightness 1 contrast 2 saturation 3 hue, -1 none);
 *   f32 factor[4] (indexed by op); i32 hue_shift (uint8(int32(hue*255))), gray, blur_r (< 0: no blur);
 *   u32 blur_ww, blur_fw (Pillow box-blur fixed-point weights); i32 n_erase; i32 ei[3], ej[3], eh[3], ew[3]; u32 seed; i32 pad.
 * lsum_ws: device uint64 [N,4] scratch. max_pixels = max h*w, max_erase = max n_erase over the batch (+ 256 when some
 * image has blur_r > 2: the three box passes per direction then run unfused for it). */
int ut2_strong_augment_u8(const void* table, int N, int max_pixels, int max_erase, unsigned long long* lsum_ws, void* stream);

/* weak augmentation (dataset_mapper.py:88-91 -> [D2] ResizeShortestEdge + RandomFlip): Pillow Image.resize(BILINEAR) bit for
 * bit (22-bit fixed-point separable resampling, uint8 between the passes) + optional horizontal flip; src uint8 [h,w,3] HWC
 * -> dst uint8 [3,new_h,new_w] CHW. tmp: uint8 [h,new_w,3]; ws: scratch of ut2_resize_workspace_bytes(h, w, new_h, new_w). */
long long ut2_resize_workspace_bytes(int h, int w, int new_h, int new_w);
int ut2_resize_flip_u8(const void* src_hwc, int h, int w, void* dst_chw, int new_h, int new_w, int flip, void* tmp_hwc, void* ws,
                       long long ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UT2_H */
