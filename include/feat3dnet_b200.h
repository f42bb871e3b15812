/*
 * feat3dnet_b200.h -- C ABI of the B200-native 3DFeat-Net sample-and-group + set-abstraction path.
 *
 * This is the drop-in boundary.  Every entry point below replaces one launcher (or one group of
 * TensorFlow graph ops) of cwlroda/3DFeatNet; the reference interface it replaces is cited as
 * file:line relative to the reference tree.  Conventions (all functions):
 *   - plain C: raw DEVICE pointers + int sizes, row-major contiguous fp32 / int32 tensors with the
 *     reference's shapes; no torch / TF types;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); the callee only
 *     enqueues work on it: it never allocates, never synchronises, never touches another stream;
 *   - zero-fills that the reference leaves to its caller (tf_sampling.cpp:174, tf_grouping.cpp:270)
 *     happen inside the callee;
 *   - return value: 0 on success, F3D_ERR_* (negative) for invalid arguments, or a positive
 *     cudaError_t from the launch.  f3d_last_error_string() describes the last non-zero code.
 *   - workspaces are caller-provided; f3d_*_workspace_bytes() tells the size.
 */
#ifndef FEAT3DNET_B200_H_
#define FEAT3DNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define F3D_ERR_INVALID_ARGUMENT (-1)
#define F3D_ERR_WORKSPACE_TOO_SMALL (-2)
#define F3D_ERR_UNSUPPORTED (-3)

/* library / device introspection */
int f3d_version(void);
const char *f3d_last_error_string(void);
/* number of kernel launches issued through this library (all threads of the process) since the last reset */
long long f3d_launch_count(void);
void f3d_reset_launch_count(void);

/* ---------------------------------------------------------------- tf_ops/sampling ------------- */

/* farthestpointsamplingLauncher(b,n,m,inp,temp,out)  tf_sampling_g.cu:203-205, op tf_sampling.cpp:95-123.
 * inp (b,n,3) f32 -> out (b,m) i32.  `temp` (the reference's 32*n float scratch) is accepted for signature
 * compatibility and ignored (may be NULL) for n <= 131072: running distances live in registers (one CTA per cloud up
 * to 16384 points, a cluster of 2/4/8 CTAs beyond).  Only n > 131072 uses it, as a (b,n) float scratch. */
int f3d_farthest_point_sample(int b, int n, int m, const float *inp, float *temp, int *out, void *stream);
/* The same plus gather_point of the samples (sample_points, models/pointnet_common.py:14-29) in the same launch:
 * new_xyz (b,m,3) = inp[b, out[b,j], :] -- identical values to f3d_gather_point on the returned indices. */
int f3d_farthest_point_sample_gather(int b, int n, int m, const float *inp, float *temp, int *out, float *new_xyz, void *stream);
/* The same on at most max_ctas CTAs (0 = one per cloud): for n <= 16384 each CTA then walks ceil(b / max_ctas) clouds one after the
 * other, so the sampling of a batch holds a chosen number of SMs (for longer) and the rest stay free for other streams.
 * Identical results for every max_ctas. */
int f3d_farthest_point_sample_gather_ctas(int b, int n, int m, const float *inp, float *temp, int *out, float *new_xyz, int max_ctas,
                                          void *stream);

/* gatherpointLauncher(b,n,m,inp,idx,out)  tf_sampling_g.cu:206-208, op tf_sampling.cpp:126-148. */
int f3d_gather_point(int b, int n, int m, const float *inp, const int *idx, float *out, void *stream);

/* scatteraddpointLauncher(b,n,m,out_g,idx,inp_g)  tf_sampling_g.cu:209-211, op tf_sampling.cpp:151-178.
 * Deterministic (no atomics): contributions to one point are added in ascending j.
 * workspace: f3d_scatter_workspace_bytes(b*m) bytes. */
int f3d_gather_point_grad(int b, int n, int m, const float *out_g, const int *idx, float *inp_g, void *workspace,
                          size_t workspace_bytes, void *stream);

/* cumsumLauncher(b,n,inp,out)  tf_sampling_g.cu:194-196.  Bit-exact: same summation DAG as the reference's blocked scan. */
int f3d_cumsum(int b, int n, const float *inp, float *out, void *stream);
/* probsampleLauncher(b,n,m,inp_p,inp_r,temp,out)  tf_sampling_g.cu:197-201, op tf_sampling.cpp:66-91.
 * temp: (b,n) f32 working space (receives the cumsum). */
int f3d_prob_sample(int b, int n, int m, const float *inp_p, const float *inp_r, float *temp, int *out, void *stream);

/* ---------------------------------------------------------------- tf_ops/grouping ------------- */

/* queryBallPointLauncher(b,n,m,radius,nsample,xyz1,xyz2,idx,pts_cnt)  tf_grouping_g.cu:179-182,
 * op tf_grouping.cpp:86-125.  xyz1 (b,n,3), xyz2 (b,m,3) -> idx (b,m,nsample) i32, pts_cnt (b,m) i32.
 * Bit-exact with the reference including the pad-with-first-hit rule and the carried "nearest"
 * fallback of empty balls (tf_grouping_g.cu:13-14,43-47). */
int f3d_query_ball_point(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2, int *idx,
                         int *pts_cnt, void *stream);

/* Same operator with a caller-provided workspace of f3d_query_ball_point_workspace_bytes(b,n) bytes: the cloud is
 * binned into an xy grid of cell size ~radius and each centre only tests its 3x3 cell neighbourhood; identical
 * results (hits are re-ordered by index through a shared-memory bitmap).  Falls back to the scan above when the
 * workspace is NULL / too small.  With m >= 4096 centres per cloud (the attention pass of inference.py:99-131 scores every point) and a
 * workspace of align256(..._workspace_bytes(b,n)) + ..._workspace_bytes(b,m) bytes the centres are binned too and visited in that
 * order, so that the warps of a CTA share their candidate cells (same rows, each written at its centre's own index).
 * Clouds of n >= 32768 points are binned per index window (<= 32 windows of consecutive indices, one cell table each; the workspace size
 * accounts for it): a centre walks its windows in ascending order and stops at nsample hits -- the reference's break at cnt == nsample,
 * tf_grouping_g.cu:31-46 -- instead of testing every point of its 3x3 cells (thousands in a KITTI-shape scan).  When xyz2 points INTO
 * xyz1 (one cloud, the centres are points of the cloud itself, as in inference.py:118-131) the centres are visited in the order of the
 * cloud's binning; the rows written are the same. */
size_t f3d_query_ball_point_workspace_bytes(int b, int n);
int f3d_query_ball_point_ws(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2, int *idx,
                            int *pts_cnt, void *workspace, size_t workspace_bytes, void *stream);
/* The two halves of f3d_query_ball_point_ws as separate calls: the grid build depends on the cloud only, so it can run
 * on a second stream while the centres (farthest_point_sample + gather_point) are still being computed.  Both need
 * the workspace (no fallback: F3D_ERR_WORKSPACE_TOO_SMALL); the query must see the same b, n, radius, xyz1, workspace. */
int f3d_ball_grid_build(int b, int n, float radius, const float *xyz1, void *workspace, size_t workspace_bytes, void *stream);
int f3d_ball_grid_query(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2, int *idx,
                        int *pts_cnt, void *workspace, size_t workspace_bytes, void *stream);

/* queryBallPoint2Launcher(b,n,m,nsample,xyz1,xyz2,radii,idx,pts_cnt)  tf_grouping_g.cu:183-186,
 * op tf_grouping.cpp:128-172.  radii (b,m).  Rows of empty balls are left untouched, as in the reference. */
int f3d_query_ball_point2(int b, int n, int m, int nsample, const float *xyz1, const float *xyz2, const float *radii,
                          int *idx, int *pts_cnt, void *stream);

/* selectionSortLauncher(b,n,m,k,dist,outi,out)  tf_grouping_g.cu:187-190, op tf_grouping.cpp:175-205.
 * dist (b,m,n) -> outi (b,m,n) i32, out (b,m,n) f32; the first k entries of each row are the k smallest
 * in the reference's (unstable, swap-based) order; the remaining n-k entries hold the leftover elements
 * exactly as the reference leaves them. */
int f3d_selection_sort(int b, int n, int m, int k, const float *dist, int *outi, float *out, void *stream);

/* knn_point(k, xyz1, xyz2)  tf_grouping.py:63-88 (TF tile/sub/square/reduce_sum + SelectionSort + slice),
 * fused: no (b,m,n) matrix in HBM.  xyz1 (b,n,c), xyz2 (b,m,c) -> val (b,m,k) squared L2, idx (b,m,k).
 * workspace: f3d_knn_workspace_bytes(b,n,m,c,k). */
size_t f3d_knn_workspace_bytes(int b, int n, int m, int c, int k);
int f3d_knn_point(int b, int n, int m, int c, int k, const float *xyz1, const float *xyz2, float *val, int *idx,
                  void *workspace, size_t workspace_bytes, void *stream);

/* The neighbourhoods in their local frames, models/pointnet_common.py:42-54 (query_and_group_points: clockwise = 1) and :104-119
 * (sample_and_group: clockwise = 0), one launch instead of group_point + tile + sub + div + cos / sin / stack (or matmul with R):
 * out (b,m,s,3) = rotate_z((xyz[idx] - centre) [/ radius when normalize_radius], angles[b,m]); angles NULL = no rotation;
 * before (b,m,s,3) or NULL receives the rows before the rotation (end_points['grouped_xyz_before']), rotation (b,m,3,3) or NULL the
 * matrices R = ((c, s, 0), (-s, c, 0), (0, 0, 1)) of :112-117 (end_points['rotation']).  Same roundings as the op chain.
 * ..._angle_grad: dangle (b,m) = gradient with respect to the angles given out (the forward's result) and gout = dL/dout -- what the
 * training step needs (the descriptor's clusters turn with the detector's orientation, models/feat3dnet.py:301-305); the gradients
 * with respect to xyz / centres go through f3d_group_point_grad on the op-by-op statement. */
int f3d_group_local_frames(int b, int n, int m, int s, const float *xyz, const float *centres, const int *idx, const float *angles,
                           int clockwise, float radius, int normalize_radius, float *before, float *rotation, float *out, void *stream);
int f3d_group_local_frames_angle_grad(int b, int m, int s, const float *out, const float *gout, int clockwise, float *dangle, void *stream);

/* groupPointLauncher(b,n,c,m,nsample,points,idx,out)  tf_grouping_g.cu:191-194, op tf_grouping.cpp:209-237. */
int f3d_group_point(int b, int n, int c, int m, int nsample, const float *points, const int *idx, float *out,
                    void *stream);

/* groupPointGradLauncher(b,n,c,m,nsample,grad_out,idx,grad_points)  tf_grouping_g.cu:195-199,
 * op tf_grouping.cpp:240-274.  Deterministic segmented reduction (no atomics): the contributions to one
 * point are added in ascending (j,k) order -- the order of the reference's CPU statement
 * (tf_ops/grouping/test/query_ball_point.cpp:68-84).
 * workspace: f3d_scatter_workspace_bytes(b*m*nsample). */
size_t f3d_scatter_workspace_bytes(long long num_slots);
/* Recommended size for b clouds of n points and slots_per_cloud slots each (>= the minimum above): also holds the offset
 * table of the per-cloud path (n <= 16384, <= 65535 slots per cloud) when the clouds have more points than slots. */
size_t f3d_scatter_add_workspace_bytes(int b, int n, long long slots_per_cloud);
int f3d_group_point_grad(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx,
                         float *grad_points, void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------- models/ (fused forward) ----- */

/* Folded eval-mode weights of the detector + descriptor (models/feat3dnet.py:277-310, layers.py:11-46):
 * every conv+BN pair is folded on the host into W' (Cin,Cout) and b' (Cout).  `packed` is one device
 * buffer of f3d_packed_weights_floats(feature_dim) floats; block i (see weights_layout.h: WeightSlot, 22
 * blocks: W,b of detection/conv0..2, conv_post_0..1, attention, orientation, description/layer1/conv0..1,
 * conv_mid_0, conv_post_0) starts at offsets[i] and holds sizes[i] floats. */
size_t f3d_packed_weights_floats(int feature_dim);
int f3d_packed_weights_num_blocks(void);
int f3d_packed_weights_offsets(int feature_dim, int *offsets, int *sizes);

/* workspace of the two fused forward entry points below (pooled per-cluster vectors) */
size_t f3d_forward_workspace_bytes(int b, int m, int feature_dim);

/* feature_detection_module forward (feat3dnet.py:90-151) on given cluster centres, eval-mode BN:
 * xyz (b,n,3), new_xyz (b,m,3), idx (b,m,nsample) from f3d_query_ball_point ->
 * attention (b,m), orientation (b,m).  Fuses group_point, translate, /radius, the 3->64->128->256 shared
 * MLP, max-pool, 256->128->64, softplus / l2norm+atan2 heads: no grouped tensor touches HBM.
 * precision: 0 = fp32 CUDA-core FFMA (exact fp32 reference path);
 *            2 = tcgen05 tensor cores, "bf16x3": operands split x = hi + lo in bf16, hi*hi + hi*lo + lo*hi accumulated
 *                in fp32 (TMEM) -- ~1e-5 relative, needs nsample == 64.
 * nsample must be 8, 16, 32, 64 or 128.
 * F3D_PRECISION_IMAGES_CACHED may be OR-ed into `precision`: the tensor-core kernels read the weights as bf16 hi/lo operand
 * images that each call builds from `packed` into fixed slots of the workspace; with the flag the caller promises that the
 * workspace still holds the images of an earlier call of the SAME function with unchanged `packed` contents and feature_dim,
 * and the image-build kernels are skipped (weights are constants between checkpoints; results are bit-identical). */
#define F3D_PRECISION_IMAGES_CACHED 0x100
/* F3D_PRECISION_SM_LIMIT(n) may be OR-ed into `precision` as well (n in 1..255, 0 = no limit): the persistent tensor-core kernels of the
 * call (one CTA per SM) launch at most n CTAs, leaving the other SMs to work the caller runs beside it on another stream -- the
 * pipeline samples the next batch (f3d_farthest_point_sample_gather_ctas) there.  Results do not depend on n. */
#define F3D_PRECISION_SM_LIMIT(n) (((n) & 0xff) << 16)
int f3d_detector_forward(int b, int n, int m, int nsample, float radius, const float *xyz, const float *new_xyz,
                         const int *idx, const float *packed, float *attention, float *orientation, int precision,
                         void *workspace, size_t workspace_bytes, void *stream);

/* feature_extraction_module forward (feat3dnet.py:154-187 -> pointnet_sa_module :9-87 ->
 * sample_and_group pointnet_common.py:69-135) on given keypoints and orientations (NULL = NoRegress):
 * -> features (b,m,feature_dim), l2-normalised.  feature_dim in {16,32,64,128} (inference.py:41). */
int f3d_descriptor_forward(int b, int n, int m, int nsample, float radius, int feature_dim, const float *xyz,
                           const float *new_xyz, const int *idx, const float *orientation, const float *packed,
                           float *features, int precision, void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------- inference.py nms ------------ */

/* nms(xyz, attention)  inference.py:226-261 for a batch of clouds, on the device (the reference: CPU, scikit-learn
 * BallTree).  xyz (b,n,3), attention (b,n) -> out_idx (b,max_keypoints) i32 indices into the cloud, out_xyz
 * (b,max_keypoints,3), out_attention (b,max_keypoints), num_keypoints (b) i32; rows beyond num_keypoints repeat the
 * strongest keypoint.  nms_radius / min_response_ratio are doubles like the reference's Python floats
 * (defaults 0.5 / 1e-2, max_keypoints 1024, num_neighbors 50: inference.py:40-47,236).  n >= num_neighbors.
 * workspace: f3d_nms_workspace_bytes(b,n). */
size_t f3d_nms_workspace_bytes(int b, int n);
int f3d_nms(int b, int n, const float *xyz, const float *attention, double nms_radius, double min_response_ratio,
            int max_keypoints, int num_neighbors, int *out_idx, float *out_xyz, float *out_attention,
            int *num_keypoints, void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------- training step --------------- */

/* conv2d(..., bn=True, is_training=True)  models/layers.py:11-46 + batch_norm_template :225-272, 1x1 kernels, channels
 * last: x (rows,cin), W (cin,cout), bias/gamma/beta (cout) -> z = x W + bias (rows,cout), batch moments mean/var (cout,
 * population variance, what the caller feeds the EMA update), y = [relu](gamma (z-mean) rsqrt(var+eps) + beta).
 * group_bias (rows/group_s, cout; NULL = none) is added to every row of its group of group_s consecutive rows: with it
 * the concat([h, tile(pooled)]) input of pointnet_sa_module's conv_mid (feat3dnet.py:60-69) never has to be built --
 * z = h W_top + (pooled W_bottom)[group] -- and its gradient comes back as dgroup_bias = per-group row sums of dz.
 * cout a multiple of 16.  precision: 0 = fp32 FFMA contractions, 2 = tcgen05 tensor cores with the bf16 hi/lo split
 * ("bf16x3", ~1e-5 relative; cin <= 256).  workspace: f3d_conv_bn_train_workspace_bytes(rows,cin,cout) (covers forward
 * and backward). */
size_t f3d_conv_bn_train_workspace_bytes(long long rows, int cin, int cout);
int f3d_conv_bn_train_forward(long long rows, int cin, int cout, const float *x, const float *W, const float *bias,
                              const float *group_bias, int group_s, const float *gamma, const float *beta, int relu, float eps,
                              float *z, float *y, float *mean, float *var, int precision, void *workspace,
                              size_t workspace_bytes, void *stream);
/* The same layer when its activation only feeds tf.reduce_max over groups of pool_s consecutive rows (the sample axis;
 * feat3dnet.py:60,75,130): pooled (rows/pool_s, cout) and inv_ties (1 / number of rows attaining the maximum) come back
 * instead of y, which is never materialised (bit-identical to the forward above followed by f3d_maxpool_samples_forward). */
int f3d_conv_bn_train_forward_pooled(long long rows, int cin, int cout, const float *x, const float *W, const float *bias,
                                     const float *group_bias, int group_s, const float *gamma, const float *beta, int relu, float eps,
                                     float *z, int pool_s, float *pooled, float *inv_ties, float *mean, float *var, int precision,
                                     void *workspace, size_t workspace_bytes, void *stream);
/* The gradient TensorFlow derives for the layer above: gy = dL/dy (rows,cout) -> dx (rows,cin; NULL = not needed),
 * dW (cin,cout), db, dgamma, dbeta (cout), through the batch statistics.  cout a power of two in 16..1024; dx needs
 * cin == 3 or a multiple of 16.  pool_s > 0: the layer's output only feeds the max-pool over groups of pool_s consecutive
 * rows; gy is then the gradient of the POOLED tensor (rows/pool_s, cout), pooled / inv_ties are the outputs of
 * f3d_maxpool_samples_forward, and the dense (rows, cout) gradient is formed on the fly instead of passing through HBM
 * (the channel reductions then run over the pooled tensors only).  y may be NULL: the activation is recomputed from z
 * with the forward's exact roundings instead of being read back.
 * Fixed-order reductions: bit-reproducible. */
int f3d_conv_bn_train_backward(long long rows, int cin, int cout, const float *x, const float *W, const float *gamma,
                               const float *beta, const float *z, const float *y, const float *mean, const float *var, int relu, float eps,
                               const float *gy, int pool_s, const float *pooled, const float *inv_ties, float *dx, float *dW,
                               float *db, float *dgamma, float *dbeta, float *dgroup_bias, int group_s, int precision,
                               void *workspace, size_t workspace_bytes, void *stream);

/* The same layer inside a CHAIN of conv2d layers (the per-point MLPs of feat3dnet.py:43-47,119-123: conv0 -> conv1 -> conv2) whose
 * intermediate activations are never written to HBM.
 *   x_coef != NULL: x is the PREVIOUS layer's pre-BN tensor z and x_coef [2][cin] its BN scale / shift (that layer's coef_out); the input
 *   rows [relu](x * scale + shift) are formed inside the contractions with the forward's exact roundings (tensor-core path, cin % 8 == 0,
 *   cin <= 128 for the weight gradient).
 *   Forward results, any subset (at least one): y (rows, cout); pool_s > 0: pooled / inv_ties (rows / pool_s, cout); coef_out [2][cout].
 *   Backward gradients: gy (dense, rows x cout; NULL = none) and / or the pooled set (pool_s, pooled, gpool = gradient of the pooled
 *   tensor, inv_ties).  With both, the activation fed the pool and a dense consumer (descriptor conv1, feat3dnet.py:47,60-66) and the two
 *   gradients are summed on the fly: neither a max-pool backward pass nor an add over (rows, cout) is run.
 * Results carry the same bits as the unchained calls on a materialised activation. */
int f3d_conv_bn_train_forward_chain(long long rows, int cin, int cout, const float *x, const float *x_coef, int x_relu, const float *W,
                                    const float *bias, const float *group_bias, int group_s, const float *gamma, const float *beta, int relu,
                                    float eps, float *z, float *y, int pool_s, float *pooled, float *inv_ties, float *coef_out, float *mean,
                                    float *var, int precision, void *workspace, size_t workspace_bytes, void *stream);
int f3d_conv_bn_train_backward_chain(long long rows, int cin, int cout, const float *x, const float *x_coef, int x_relu, const float *W,
                                     const float *gamma, const float *beta, const float *z, const float *mean, const float *var, int relu,
                                     float eps, const float *gy, int pool_s, const float *pooled, const float *gpool, const float *inv_ties,
                                     float *dx, float *dW, float *db, float *dgamma, float *dbeta, float *dgroup_bias, int group_s,
                                     int precision, void *workspace, size_t workspace_bytes, void *stream);

/* out (rows, nout) = x (rows, k) W with W (k, nout) row-major, and its gradients dx = g W^T, dW = x^T g, on the tensor-core contractions of the
 * training layers (forward: 3-way bf16 split, fp32-grade; gradients: 2-way split).  Used for the per-cluster term of conv_mid: the pooled half of
 * concat([h, tile(max h)]) (models/feat3dnet.py:60-69) contributes pooled W_bottom once per cluster.  k <= 256; dW needs k % 8 == 0, k <= 128,
 * nout % 16 == 0, nout <= 256.  dx / dW may be NULL.  workspace: f3d_linear_workspace_bytes(rows, k, nout). */
size_t f3d_linear_workspace_bytes(long long rows, int k, int nout);
int f3d_linear_forward(long long rows, int k, int nout, const float *x, const float *W, float *out, void *workspace, size_t workspace_bytes, void *stream);
int f3d_linear_backward(long long rows, int k, int nout, const float *x, const float *W, const float *g, float *dx, float *dW, void *workspace,
                        size_t workspace_bytes, void *stream);

/* tf.reduce_max(new_points, axis=[2])  models/feat3dnet.py:138,147,182 on a channels-last (groups, s, c) tensor
 * (c % 4 == 0) -> out (groups, c) and inv_ties (groups, c; NULL = not wanted) = 1 / (number of samples attaining the
 * maximum); and its gradient: samples attaining the maximum share gout equally (TF _MinOrMaxGrad). */
int f3d_maxpool_samples_forward(long long groups, int s, int c, const float *x, float *out, float *inv_ties, void *stream);
int f3d_maxpool_samples_backward(long long groups, int s, int c, const float *x, const float *out, const float *inv_ties,
                                 const float *gout, float *dx, void *stream);

/* Feat3dNet.get_loss  models/feat3dnet.py:315-357 (+ pairwise_dist, models/layers.py:49-62): attention-weighted
 * triplet loss over anchor / positive / negative descriptors fa, fp, fn (b,m,f) and anchor attention att (b,m; NULL =
 * uniform 1/m weights, the Attention=False branch).  loss: 1 float.  When dfa, dfp, dfn are non-NULL the same call
 * also writes the gradients of the loss w.r.t. the descriptors (b,m,f) and, if datt is non-NULL, w.r.t. att (b,m);
 * ties of the row minima share the gradient equally like tf.reduce_min.  Deterministic (no atomics).
 * workspace: f3d_triplet_loss_workspace_bytes(b,m). */
size_t f3d_triplet_loss_workspace_bytes(int b, int m);
int f3d_triplet_loss(int b, int m, int f, float margin, const float *fa, const float *fp, const float *fn,
                     const float *att, float *loss, float *dfa, float *dfp, float *dfn, float *datt, void *workspace,
                     size_t workspace_bytes, void *stream);

/* The detector's per-cluster heads in training (models/feat3dnet.py:142-149): h (rows,k) -> attention = softplus(h w_att + b_att)
 * (rows), orientation = atan2 of the l2-normalised (h w_ori + b_ori) (rows); w_att (k), w_ori (k,2).  k in {32,64,96,128}.
 * Backward: g_att / g_ori = dL/d attention, dL/d orientation (either may be NULL) -> dh (rows,k; NULL to skip), dw_att (k),
 * db_att (1), dw_ori (k,2), db_ori (2); deterministic (fixed-order partial sums).  workspace: f3d_detector_heads_workspace_bytes(k). */
size_t f3d_detector_heads_workspace_bytes(int k);
int f3d_detector_heads_forward(long long rows, int k, const float *h, const float *w_att, const float *b_att, const float *w_ori,
                               const float *b_ori, float *attention, float *orientation, void *stream);
int f3d_detector_heads_backward(long long rows, int k, const float *h, const float *w_att, const float *b_att, const float *w_ori,
                                const float *b_ori, const float *g_att, const float *g_ori, float *dh, float *dw_att, float *db_att,
                                float *dw_ori, float *db_ori, void *workspace, size_t workspace_bytes, void *stream);

/* Feat3dNet.get_train_op  models/feat3dnet.py:359-375: tf.train.AdamOptimizer(lr).minimize over all variables in
 * ONE launch.  records: device array of num_records x {float *param; const float *grad; float *m; float *v;
 * long long n;} (40 bytes each); max_n = the largest n.  step = 1-based update count; grad_scale multiplies every
 * gradient first (1/world_size after a sum all-reduce).  theta -= lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps).
 * step_dev (device int64 or NULL): when given, t = *step_dev + 1 is read on the device and the counter is advanced
 * afterwards (`step` is ignored) -- the form a captured CUDA graph of the training step replays. */
int f3d_adam_step(int num_records, const void *records, long long max_n, float lr, float beta1, float beta2, float eps,
                  long long step, float grad_scale, long long *step_dev, void *stream);

/* One output row [x y z | attention | orientation | descriptor] per keypoint (rows x (5 + feature_dim) floats): the packed
 * result the end-to-end path copies to the host in one transfer; inference.py:174-177 writes its [xyz | descriptor] file
 * rows from the same fields. */
int f3d_pack_rows(long long rows, int feature_dim, const float *xyz, const float *attention, const float *orientation,
                  const float *features, float *out, void *stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Registration: the MATLAB evaluation step of the reference on the device.
 *
 * scripts/computeAndVisualizeMatches.m:43-44  [~, m] = pdist2(desc2, desc1, 'euclidean', 'smallest', 1):
 * match[i] = index of the row of desc2 (n2 x dim) nearest to row i of desc1 (n1 x dim), lowest index on ties;
 * dist2 (optional) receives the squared distance. */
int f3d_match_descriptors(int n1, int n2, int dim, const float *desc1, const float *desc2, int *match, float *dist2, void *stream);
/* scripts/external/ransacfitRt.m + ransac.m + estimateRigidTransform.m: pts1[k] ~ R pts2[k] + t over npts correspondences
 * (npts x 3 each).  triples (ntrials x 3) are the 3-point samples of the trials IN ORDER (MATLAB's randsample stream is
 * not reproducible, so the caller draws them); every trial is scored in parallel and the reference's sequential
 * bookkeeping (">=" best-score rule, adaptive N at p = 0.99, max_trials, default 10000) is replayed over the scores.
 * Rt: 12 doubles, row-major 3x4, refit on the inliers (NaN if fewer than 3); inlier_mask: npts bytes;
 * info: {inliers of the chosen trial, trialcount, chosen trial, status: 1 = triples exhausted before the stop rule}. */
size_t f3d_ransac_workspace_bytes(int npts, int ntrials);
int f3d_ransac_fit_rt(int npts, const float *pts1, const float *pts2, int ntrials, const int *triples, float threshold,
                      int max_trials, double *Rt, unsigned char *inlier_mask, int *info, void *workspace, size_t workspace_bytes,
                      void *stream);
/* estimateRt.m / estimateRigidTransform.m on the correspondences with mask[k] != 0 (all when mask is NULL); eps (optional)
 * receives the smallest singular value S(4,4). */
int f3d_rigid_fit(int npts, const float *pts1, const float *pts2, const unsigned char *mask, double *Rt, double *eps, void *stream);

/* ---------------------------------------------------------------- bring-up / debugging ---------- */

/* Single-CTA tcgen05 self test: D[128 x N] = A[128 x K] * B[N x K]^T from canonical K-major no-swizzle bf16 operand
 * images (element (r,k) at (k/8)*lbo + (r/8)*sbo + (r%8)*16 + (k%8)*2 bytes).  a_in_tmem != 0 first copies A into
 * tensor memory (tcgen05.cp) and feeds the MMA from there.  Not part of the reference surface. */
int f3d_debug_umma_selftest(const void *a_img, const void *b_img, float *D, int N, int K, int lbo_a, int sbo_a, int lbo_b,
                            int sbo_b, int a_bytes, int b_bytes, int a_in_tmem, void *stream);
/* Bring-up: the tensor-core weight-gradient contraction alone (partW: 2*SMs x cin x cout floats of per-CTA partials);
 * dbg bit 0 skips the operand staging (TMA + conversion), bit 1 the MMAs, bit 2 the conversion's shared-memory loads, bit 3 its stores,
 * bit 4 the conversion altogether (micro-benchmarking: tools/wgrad_tc_phases.py). */
int f3d_debug_wgrad_tc(long long rows, int cin, int cout, const float *x, const float *dz, float *partW, int dbg, void *stream);
/* Test / measurement aid: 0 = pool-only training layers materialise dz (bn_bwd_apply) for wgrad and dgrad, 1 (default) = dz is formed
 * inside the two contractions from z and the pooled tensors.  Both paths give the same dW / dx bits.  Returns the previous value. */
int f3d_debug_set_fuse_dz(int on);
/* Test / measurement aid: 0 = pool-only training layers take the pooled maximum / tie counts with a pass over z (bn_apply_pool), 1 (default) =
 * from statistics gathered in the forward contraction's epilogue.  Same bits.  Returns the previous value. */
int f3d_debug_set_epilogue_pool(int on);
/* Measurement aid: 0 = the chunk-owning streaming kernels of the training layers (BN-backward passes, xyz-layer forward) give every block one
 * contiguous 1/grid of the rows, 1 (default) = small chunks round-robin over the grid (4.4 -> 5.4 TB/s: 592 equal streams in lockstep hit the
 * same DRAM channels).  The summation order differs between the two; each is deterministic. */
int f3d_debug_set_row_walk(int on);
/* Measurement aids: skip phases of the lin_tc kernels (bit 0 operand conversion, 1 MMAs, 2 epilogue stores, 3 TMA fetches; 0 = run all;
 * results are garbage with a bit set), and the contraction alone: out (rows, nout) = x (rows, k) W^T, W (nout, k) row-major, nsplit 2 | 3,
 * part = NULL or 2 * 2 * nout floats per row CTA of column-sum partials, wimg = f3d_debug_lin_tc_weight_bytes(k, nout) of scratch. */
int f3d_debug_set_lin_tc_phases(int skip_mask);
/* Measurement aid: the smallest (padded) K for which lin_tc launches its warp-specialised kernel (default 128).  Returns the previous value. */
int f3d_debug_set_lin_tc_pipe_min_k(int k);
/* buf: device memory of 2 * 64 * 16 long long receiving clock64() stamps of threads 0 and 255 of CTA (0,0) of lin_tc_kernel (16 slots per
 * tile, first 64 tiles); NULL switches the trace off. */
int f3d_debug_lin_tc_trace(void *buf);
size_t f3d_debug_lin_tc_weight_bytes(int k, int nout);
int f3d_debug_lin_tc(long long rows, int k, int nout, const float *x, const float *W, float *out, float *part, void *wimg, int nsplit, void *stream);
size_t f3d_detector_tc_weight_bytes(void);
/* Measurement aid: `groups` x `per_group` back-to-back tcgen05.mma of shape (128 * cta_group) x N x 16 (A from tensor memory, B a
 * zero-filled bf16 no-swizzle image in shared memory, K- or MN-major), one commit per group, on `ctas` CTAs at once (cta_group 2:
 * clusters of 2, M = 256 across the pair).  out: 2 int64 per CTA = {clock64 cycles of the loop, instructions}.  */
int f3d_debug_umma_bench(int cta_group, int N, int groups, int per_group, int b_mn_major, int ctas, void *out, void *stream);
/* Bring-up: device buffer of (tiles per CTA) x 16 int64 receiving CTA 0's clock64() timeline of the detector tensor
 * kernel (slots: 0/1/2 MMA warp, 4-6 producer, 8-13 epilogue); NULL disables. */
void f3d_debug_set_timeline(void *buf);
void f3d_debug_set_timeline_desc(void *buf); /* same for the descriptor tensor kernel */
/* Measurement aid: while enabled (enable != 0 also clears earlier records) the hot kernels -- fps_group, bq_grid_query, det_rows_tc,
 * desc_rows_tc, post_tc, lin_tc, wgrad_tc, the BN passes -- are bracketed by CUDA events on the stream they are launched on (eager
 * launches only, not under graph capture).  f3d_debug_kernel_timings waits for them and returns their count (<= max): names
 * (max x 48 chars), durations (ms) and the ALGORITHMIC bytes or flops of each launch as its call site states them. */
void f3d_debug_kernel_timer(int enable);
int f3d_debug_kernel_timings(int max, char *names, float *ms, double *units);

#ifdef __cplusplus
}
#endif
#endif /* FEAT3DNET_B200_H_ */
