/*
 * scn_b200.h - C-ABI of libscn_b200.so, the B200-native replacement for the SparseConvNet
 * "SCN" extension on the sparse3d-backbone hot path.
 *
 * Every entry point is what the reference's pybind surface for this path would bind
 * (reference: SparseConvNet/sparseconvnet/SCN/pybind.cpp, sparseconvnet.h,
 * sparseconvnet_cuda.cpp).  Differences forced by a C ABI:
 *   - no tensor types: raw DEVICE pointers (fp32 features / weights), plain sizes, and the
 *     CUDA stream as a void* (cudaStream_t);
 *   - the reference passes an empty output tensor and resize_()s it inside C++
 *     (e.g. CUDA/Convolution.cpp:38); here the caller first asks for the row count
 *     (scn_get_nactive / the *_prepare calls) and allocates the output itself;
 *   - errors: every call returns 0 on success, non-zero on failure, and scn_last_error()
 *     returns a thread-local message (reference: C asserts / ATen exceptions).
 *
 * Conventions: feature matrices are row-major [rows, planes] fp32, contiguous.
 * Weights are [K, 1, nIn, nOut] fp32 exactly like the reference Parameter
 * (submanifoldConvolution.py:24-26); K enumerates the filter box row-major with the LAST
 * spatial dimension fastest (RectangularRegions.h:31-38).  `spatial_size`, `filter_size`,
 * `filter_stride` are int64[3].  All launches go to `stream`; the only host synchronisations
 * are the documented count read-backs inside the *_prepare / first-use rulebook builds and
 * the explicit copy-out calls (scn_get_spatial_locations, scn_*_rulebook_*).
 * There is NO CPU fallback anywhere in this library.
 */
#ifndef SCN_B200_H
#define SCN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct scn_metadata scn_metadata_t;

/* Precision of the convolution contraction (accumulation is always fp32). */
enum {
  SCN_PRECISION_FP32 = 0,   /* exact fp32 FFMA tiles (parity mode)                               */
  SCN_PRECISION_TF32 = 1,   /* tf32 operands on tcgen05 tensor cores, fp32 accumulators in TMEM  */
  SCN_PRECISION_FP32_3XTF32 = 2, /* fp32 accuracy on the tensor cores: x = hi + lo, w = whi + wlo in tf32,
                                   hi*whi + lo*whi + hi*wlo accumulated in fp32 (error ~2^-20 relative);
                                   shapes the tensor path does not take use the exact FFMA tiles           */
  SCN_PRECISION_BF16 = 3       /* forward / input-gradient contractions with bf16 operands (features rounded to
                                   nearest, kind::f16 MMAs, A operand from tensor memory), fp32 accumulation;
                                   the weight gradient runs in single-pass tf32                              */
};

/* ---- library ---------------------------------------------------------------------- */
const char *scn_last_error(void);
int scn_version(void);
/* replaces pybind.cpp:234 n_rulebook_bits() */
int scn_n_rulebook_bits(void);

/* ---- Metadata<3> handle (replaces class Metadata_3, pybind.cpp:11-32,205;
 *      Metadata/Metadata.h:44-163) --------------------------------------------------- */
int scn_metadata_create(int dimension, scn_metadata_t **out);
void scn_metadata_destroy(scn_metadata_t *m);
/* Metadata::clear (Metadata.cpp:51-62) */
int scn_metadata_clear(scn_metadata_t *m, void *stream);
/* Metadata::getNActive (Metadata.cpp:64-67); -1 in *n_active if the scale does not exist */
int scn_get_nactive(scn_metadata_t *m, const int64_t *spatial_size, int64_t *n_active);
/* number of samples = grids[...].size() (CPU/SparseToDense.cpp:43) */
int scn_get_batch_size(scn_metadata_t *m, int64_t *batch_size);
/* Metadata::getSpatialLocations (Metadata.cpp:149-168): int64 [nActive,4] rows (x,y,z,batch)
 * written to HOST memory `out_host` (the reference returns a CPU tensor). Synchronises. */
int scn_get_spatial_locations(scn_metadata_t *m, const int64_t *spatial_size,
                              int64_t *out_host, void *stream);
/* same rows, int32 [nActive,4], left on the device (no sync) */
int scn_get_spatial_locations_device(scn_metadata_t *m, const int64_t *spatial_size,
                                     int32_t *out_dev, void *stream);

/* ---- voxelisation front end (replaces the numpy quantiser in
 *      data3d/suncg_utils/suncg_dataset.py:126-188 + collate data3d/data.py:25-37) ----
 * xyz: DEVICE float64 [n,3] metres; computes a = xyz*scale in fp64, a -= min(a) per column
 * (min taken over this sample), keeps rows with 0 <= a < full_scale, truncates to int64.
 * coords_out: DEVICE int64 [n,4] (x,y,z,batch_idx) compacted in input order; keep_out:
 * DEVICE uint8 [n] mask; n_kept read back to the host (one sync). */
int scn_quantize_points(const double *xyz, int64_t n, double scale, const int64_t *full_scale,
                        int64_t batch_idx, int64_t *coords_out, uint8_t *keep_out,
                        int64_t *n_kept, void *stream);

/* Whole-batch front end: the dataset's per-building quantisation AND the collate, on the device, fed by ONE float32
 * buffer (replaces SUNCGDataset.__getitem__, data3d/suncg_utils/suncg_dataset.py:126-188, and trainMerge,
 * data3d/data.py:25-37).  points: DEVICE float32 [n, n_cols], columns 0-2 = xyz in metres, the rest features
 * (rgb, normal), the buildings of the batch back to back; first_dev: DEVICE int64 [n_buildings] first point of
 * each building; matrix: HOST double[9], row-major, the dataset's `m` (eye(3) * scale times its zoom / flip /
 * rotation augmentations): a = xyz @ matrix in float64, a -= a.min(0) per building, rows with 0 <= a < full_scale
 * kept, coordinates truncated.  coords_out: DEVICE int64 [n,4] (x,y,z,building) and feats_out: DEVICE float32
 * [n, n_cols], both compacted in input order; xyz_feature != 0 writes a / scale into feature columns 0-2 (:160-162).
 * n_kept is read back to the host (one sync).  Results equal the numpy pipeline bit for bit when `matrix` is
 * diagonal; with rotation / zoom the float64 products are summed as (x m0 + y m1) + z m2, which numpy's BLAS
 * matmul does not pin. */
int scn_voxelize_batch(const float *points, int64_t n, int64_t n_cols, const int64_t *first_dev,
                       int64_t n_buildings, const double *matrix, double scale, const int64_t *full_scale,
                       int xyz_feature, int64_t *coords_out, float *feats_out, int64_t *n_kept, void *stream);

/* ---- InputLayer / OutputLayer (replaces InputLayer_updateOutput / _updateGradInput,
 *      OutputLayer_updateOutput / _updateGradInput: pybind.cpp:154-170,
 *      CPU/IOLayers.cpp:45-140, Metadata::inputLayer Metadata.cpp:406-417,
 *      inputLayerRules IOLayersRules.h:19-125) ---------------------------------------
 * coords: int64 [n_points, n_cols] (n_cols = 3 or 4, last column = batch index), in HOST
 * memory when coords_on_device == 0 (the reference contract: ioLayers.py:60) or DEVICE
 * memory otherwise.  Builds the scale-0 hash grid, numbers the active sites in order of
 * first occurrence, builds the point->site CSR.  One host sync (returns *n_active). */
int scn_input_layer_prepare(scn_metadata_t *m, const int64_t *spatial_size,
                            const int64_t *coords, int64_t n_points, int n_cols,
                            int coords_on_device, int64_t batch_size, int mode,
                            void *stream, int64_t *n_active);
/* B200 extension: scn_input_layer_prepare followed by n_ops rulebook builds in ONE call (the whole
 * integer work of a batch; meant for a prefetch thread - a single foreign call never re-enters the
 * host language between builds).  plan: n_ops x 13 int64 =
 * [kind (0 submanifold, 1 strided), in_spatial_size[3], out_spatial_size[3], filter[3], stride[3]]. */
int scn_build_plan(scn_metadata_t *m, const int64_t *spatial_size, const int64_t *coords,
                   int64_t n_points, int n_cols, int coords_on_device, int64_t batch_size, int mode,
                   const int64_t *plan, int n_ops, void *stream, int64_t *n_active);
/* out[site] = sum/mean/first/last of in[points of site]; in [n_points,planes], out [n_active,planes] */
int scn_input_layer_forward(scn_metadata_t *m, const float *in_feats, float *out_feats,
                            int64_t n_planes, void *stream);
/* d_in [n_points,planes] (fully written), d_out [n_active,planes] */
int scn_input_layer_backward(scn_metadata_t *m, float *d_in_feats, const float *d_out_feats,
                             int64_t n_planes, void *stream);
/* OutputLayer: out[point] = in[site(point)] (average=false, CPU/IOLayers.cpp:91-109) */
int scn_output_layer_forward(scn_metadata_t *m, const float *in_feats, float *out_feats,
                             int64_t n_planes, void *stream);
int scn_output_layer_backward(scn_metadata_t *m, float *d_in_feats, const float *d_out_feats,
                              int64_t n_planes, void *stream);
/* header of the input rulebook: [mode, maxActive, nInputRows, nOutputRows] (IOLayersRules.h:10-15) */
int scn_input_rulebook_header(scn_metadata_t *m, int64_t header[4], void *stream);
/* reference-format table nOutputRows x (1+maxActive) int32 -> HOST */
int scn_input_rulebook_copy(scn_metadata_t *m, int32_t *out_host, void *stream);

/* ---- rulebooks (replaces Metadata::getSubmanifoldRuleBook Metadata.cpp:430-444,
 *      ::getRuleBook :485-512, ::getSparseToDenseRuleBook :469-483;
 *      SubmanifoldConvolutionRules.h:13-87, ConvolutionRules.h:12-151) ----------------
 * *_prepare builds (or finds cached) the rulebook; counts_host receives the number of
 * (in,out) pairs per kernel offset (K = prod(filter_size) entries).  One host sync on first
 * build, none on a cache hit. */
int scn_submanifold_rulebook_prepare(scn_metadata_t *m, const int64_t *spatial_size,
                                     const int64_t *filter_size, void *stream,
                                     int64_t *counts_host);
/* strided convolution rulebook; creates the output grid as a side effect and returns its
 * nActive (Metadata.cpp:497-507). Deconvolution uses the same rulebook with roles swapped
 * (CPU/Deconvolution.cpp:15-16). */
int scn_conv_rulebook_prepare(scn_metadata_t *m, const int64_t *in_spatial_size,
                              const int64_t *out_spatial_size, const int64_t *filter_size,
                              const int64_t *filter_stride, void *stream,
                              int64_t *n_out_active, int64_t *counts_host);
/* copy the pairs of one offset to HOST as int32 [count,2] = (in,out) like RuleBook (Metadata.h:35) */
int scn_submanifold_rulebook_copy(scn_metadata_t *m, const int64_t *spatial_size,
                                  const int64_t *filter_size, int64_t offset,
                                  int32_t *pairs_host, void *stream);
int scn_conv_rulebook_copy(scn_metadata_t *m, const int64_t *in_spatial_size,
                           const int64_t *filter_size, const int64_t *filter_stride,
                           int64_t offset, int32_t *pairs_host, void *stream);
/* SparseToDense rules: int32 [nActive,2] = (row, linear offset (x*Y+y)*Z+z) and the per-row
 * sample index int32 [nActive] -> HOST (ConvolutionRules.h:110-128) */
int scn_sparse_to_dense_rules_copy(scn_metadata_t *m, const int64_t *spatial_size,
                                   int32_t *rules_host, int32_t *sample_host, void *stream);

/* ---- convolutions (replaces SubmanifoldConvolution_updateOutput/_backward
 *      pybind.cpp:134-143, Convolution_* :54-65, Deconvolution_* :78-89;
 *      CPU/Convolution.cpp:46-185, CPU/Deconvolution.cpp:8-77) ------------------------
 * in [nIn,n_in_planes]; out [nOut,n_out_planes] (fully written: zero + bias + sum);
 * bias may be NULL.  *macs receives sum_k pairs_k*nIn*nOut like the reference's return.
 * backward: d_in fully written; d_weight [K,1,nIn,nOut] fully written (matmul_out
 * overwrites, CPU/Convolution.cpp:110); d_bias (may be NULL) = column sums of d_out.
 * weight_tag (may be NULL): {identity token of the weight tensor, its in-place version counter}.
 * With a tag the tensor-core forward pass packs the weights' operand images for BOTH passes in one
 * launch and the dX pass of the same version reuses them; without it every call packs its own. */
int scn_submanifold_conv_forward(scn_metadata_t *m, const int64_t *spatial_size,
                                 const int64_t *filter_size, const float *in, float *out,
                                 const float *weight, const float *bias, int64_t n_in_planes,
                                 int64_t n_out_planes, int precision, void *stream, double *macs,
    const int64_t *weight_tag);
int scn_submanifold_conv_backward(scn_metadata_t *m, const int64_t *spatial_size,
                                  const int64_t *filter_size, const float *in, float *d_in,
                                  const float *d_out, const float *weight, float *d_weight,
                                  float *d_bias, int64_t n_in_planes, int64_t n_out_planes,
                                  int precision, void *stream, const int64_t *weight_tag);
int scn_conv_forward(scn_metadata_t *m, const int64_t *in_spatial_size,
                     const int64_t *out_spatial_size, const int64_t *filter_size,
                     const int64_t *filter_stride, const float *in, float *out,
                     const float *weight, const float *bias, int64_t n_in_planes,
                     int64_t n_out_planes, int precision, void *stream, double *macs,
    const int64_t *weight_tag);
int scn_conv_backward(scn_metadata_t *m, const int64_t *in_spatial_size,
                      const int64_t *out_spatial_size, const int64_t *filter_size,
                      const int64_t *filter_stride, const float *in, float *d_in,
                      const float *d_out, const float *weight, float *d_weight, float *d_bias,
                      int64_t n_in_planes, int64_t n_out_planes, int precision, void *stream, const int64_t *weight_tag);
/* in lives on the COARSE scale (in_spatial_size), out on the FINE scale (out_spatial_size) */
int scn_deconv_forward(scn_metadata_t *m, const int64_t *in_spatial_size,
                       const int64_t *out_spatial_size, const int64_t *filter_size,
                       const int64_t *filter_stride, const float *in, float *out,
                       const float *weight, const float *bias, int64_t n_in_planes,
                       int64_t n_out_planes, int precision, void *stream, double *macs,
    const int64_t *weight_tag);
int scn_deconv_backward(scn_metadata_t *m, const int64_t *in_spatial_size,
                        const int64_t *out_spatial_size, const int64_t *filter_size,
                        const int64_t *filter_stride, const float *in, float *d_in,
                        const float *d_out, const float *weight, float *d_weight, float *d_bias,
                        int64_t n_in_planes, int64_t n_out_planes, int precision, void *stream, const int64_t *weight_tag);

/* ---- NetworkInNetwork (replaces NetworkInNetwork_updateOutput / _updateGradInput /
 *      _accGradParameters, sparseconvnet.h:50-60, CPU/NetworkInNetwork.cpp:8-46) ------ */
int scn_nin_forward(const float *in, float *out, const float *weight, const float *bias,
                    int64_t n_rows, int64_t n_in_planes, int64_t n_out_planes, int precision,
                    void *stream, double *macs);
int scn_nin_backward(const float *in, float *d_in, const float *d_out, const float *weight,
                     float *d_weight, float *d_bias, int64_t n_rows, int64_t n_in_planes,
                     int64_t n_out_planes, int precision, void *stream);

/* ---- BatchNormalization fused with (Leaky)ReLU (replaces BatchNormalization_updateOutput /
 *      _backward, sparseconvnet.h:21-32, CPU/BatchNormalization.cpp:13-107) ------------
 * train == 1: batch statistics (biased var for normalisation, unbiased for running_var),
 * running stats updated in place with `momentum` weighting the OLD value.
 * train == 0: normalise with running_mean / running_var as passed.
 * train == 2: evaluation with the statistics of the current batch and the UNBIASED variance, running
 * buffers untouched - the reference's eval mode with track_running_stats=False
 * (sparseconvnet/batchNormalization.py:51-56 computes features.mean(0) / features.var(0) eagerly and
 * passes them to the train == 0 path; here the statistics kernel of the train path does it).
 * weight / bias may be NULL.  Unlike the reference, d_out is NOT modified in place. */
int scn_batchnorm_forward(const float *in, float *out, float *save_mean, float *save_invstd,
                          float *running_mean, float *running_var, const float *weight,
                          const float *bias, float eps, float momentum, int train,
                          float leakiness, int64_t n_rows, int64_t n_planes, void *stream);
int scn_batchnorm_backward(const float *in, float *d_in, const float *out, const float *d_out,
                           const float *save_mean, const float *save_invstd,
                           const float *weight, float *d_weight, float *d_bias,
                           float leakiness, int64_t n_rows, int64_t n_planes, void *stream);
/* same, with d_in = BN gradient + residual (residual [n_rows,n_planes] may be NULL or alias d_in): the
 * second gradient of a value that feeds both a BatchNorm and a skip connection, added in the same pass */
int scn_batchnorm_backward_add(const float *in, float *d_in, const float *out, const float *d_out,
                               const float *save_mean, const float *save_invstd,
                               const float *weight, float *d_weight, float *d_bias,
                               float leakiness, int64_t n_rows, int64_t n_planes,
                               const float *residual, void *stream);
/* Same, given the forward's affine parameters: with recompute_mask != 0 `weight` / `bias` are the layer's gamma /
 * beta (NULL = 1 / 0) and the (Leaky)ReLU mask is taken from the sign of the recomputed pre-activation
 * fmaf(in, invstd * gamma, fma(-mean, invstd * gamma, beta)) - bit-identical to the forward's - so `out` is not
 * read (5 instead of 7 tensor passes over [n, C]); `out` may then be NULL.  Used by the layer-graph executor. */
int scn_batchnorm_backward_fused(const float *in, float *d_in, const float *out, const float *d_out,
                                 const float *save_mean, const float *save_invstd, const float *weight,
                                 const float *bias, int recompute_mask, float *d_weight, float *d_bias,
                                 float leakiness, int64_t n, int64_t C, const float *residual, void *stream);

/* ---- SparseToDense (replaces SparseToDense_updateOutput / _updateGradInput,
 *      pybind.cpp:124-133, CPU/SparseToDense.cpp:35-87) --------------------------------
 * out: dense [batch, n_planes, X, Y, Z] fp32, fully written (zero-filled then scattered). */
int scn_sparse_to_dense_forward(scn_metadata_t *m, const int64_t *spatial_size, const float *in,
                                float *out, int64_t n_planes, int64_t batch_size, void *stream);
int scn_sparse_to_dense_backward(scn_metadata_t *m, const int64_t *spatial_size, float *d_in,
                                 const float *d_out, int64_t n_planes, int64_t batch_size,
                                 void *stream);

/* B200 extension for tools_3d_2d.py:7-48 (sparse_3d_to_dense_2d densifies all of [X, Y, Z] and slices out the occupied
 * extent): scn_grid_extent returns ext = [max x + 1, max y + 1, max z + 1, max batch index + 1] of the grid (one 16-byte
 * read-back instead of a host copy of every location), the _cropped_ calls densify straight into
 * [batch, n_planes, ext0, ext1, ext2] = dense[:, :, :ext0, :ext1, :ext2] of the full tensor. */
int scn_grid_extent(scn_metadata_t *m, const int64_t *spatial_size, int64_t ext_out[4], void *stream);
int scn_sparse_to_dense_cropped_forward(scn_metadata_t *m, const int64_t *spatial_size, const int64_t *ext,
                                        const float *in, float *out, int64_t n_planes, int64_t batch_size, void *stream);
int scn_sparse_to_dense_cropped_backward(scn_metadata_t *m, const int64_t *spatial_size, const int64_t *ext,
                                         float *d_in, const float *d_out, int64_t n_planes, int64_t batch_size,
                                         void *stream);

/* ---- layer-graph executor (B200 extension) -------------------------------------------
 * The reference runs the backbone as ~130 Python autograd Functions per direction
 * (sparseconvnet/{submanifoldConvolution,convolution,deconvolution,batchNormalization}.py, fpn_net.py:168-265);
 * at batch 1 the enqueue work of that Python layer, not the GPU, bounds the step.  A graph is the same
 * sequence of layer ops as a flat array, executed by ONE call per direction through exactly the entry
 * points above (same kernels, same results).  Values are feature matrices [rows, planes] identified by
 * index; the caller owns every buffer (torch allocates them: values, gradients, parameter gradients).
 *   kind 1 submanifold conv, 2 convolution, 3 deconvolution (in0 -> out; p0 weight, p1 bias or -1)
 *   kind 4 BatchNorm(Leaky)ReLU (in0 -> out; p0 gamma, p1 beta, p2 running_mean, p3 running_var;
 *          save_off = float offset of this op's {mean[C], invstd[C]} in bn_save)
 *   kind 5 add (out = in0 + in1) */
typedef struct scn_graph_op {
  int32_t kind, in0, in1, out;
  int32_t n_in_planes, n_out_planes;
  int32_t p0, p1, p2, p3;
  int64_t in_ss[3], out_ss[3], filter[3], stride[3];
  int64_t save_off;
  float eps, momentum, leakiness, pad_;
} scn_graph_op_t;
/* values[v]: device pointer of value v (read for inputs, fully written for outputs); rows[v] its row count.
 * params[p]: device pointer of parameter / buffer p; param_tags[2p..2p+1]: weight tag or {0,0}.
 * *macs (may be NULL) receives the summed multiply-add count of the convolutions. */
int scn_graph_forward(scn_metadata_t *m, const scn_graph_op_t *ops, int32_t n_ops, float *const *values,
                      const int64_t *rows, float *const *params, const int64_t *param_tags, float *bn_save,
                      int train, int precision, void *stream, double *macs);
/* Reverse sweep.  out_grads[v] (NULL if none): gradient arriving from outside for value v (read only).
 * grads[v]: caller-provided buffer [rows[v], planes] for the gradient of value v, NULL if that gradient is
 * not wanted (then nothing flows further back through v).  Ops whose output receives no gradient are
 * skipped.  param_grads[p]: buffer for the gradient of parameter p (fully written if the op runs);
 * param_written[p] (host) is set to 1 for every parameter gradient written.  scratch: device buffer of
 * scratch_floats >= max rows*planes over the values (used when a convolution's input gradient has to be
 * added to an existing one). */
int scn_graph_backward(scn_metadata_t *m, const scn_graph_op_t *ops, int32_t n_ops, int32_t n_values,
                       float *const *values, const int64_t *rows, float *const *params,
                       const int64_t *param_tags, const float *bn_save, float *const *grads,
                       const float *const *out_grads, float *const *param_grads, uint8_t *param_written,
                       float *scratch, int64_t scratch_floats, int precision, void *stream);

/* Same reverse sweep with progress marks for an overlapped gradient all-reduce (DistributedDataParallel's
 * bucketed reduction, tools/train_net_sparse3d.py:64-69): mark_ops in DESCENDING order; event mark_events[j]
 * (scn_event_create) is recorded on `stream` as soon as the sweep has passed op mark_ops[j], i.e. when the
 * parameter gradients of every op with index >= mark_ops[j] are final.  param_grads may point straight into a flat
 * gradient bucket: a collective on another stream that waits for event j may then reduce those ranges while the
 * sweep continues. */
int scn_graph_backward_marked(scn_metadata_t *m, const scn_graph_op_t *ops, int32_t n_ops, int32_t n_values,
                              float *const *values, const int64_t *rows, float *const *params,
                              const int64_t *param_tags, const float *bn_save, float *const *grads,
                              const float *const *out_grads, float *const *param_grads, uint8_t *param_written,
                              float *scratch, int64_t scratch_floats, int precision, void *stream, int32_t n_marks,
                              const int32_t *mark_ops, void *const *mark_events);
/* 2 (default): the reverse sweep leaves the weight gradients on the companion stream until its marks / its end, so
 * they run under the following layers; 1: companion stream with one join per layer (the per-layer calls' behaviour);
 * 0: everything on the caller's stream - no two kernels of a step overlap, which is what the per-class roofline
 * pass of bench.py wants to time */
int scn_set_graph_overlap(int mode);
/* Programmatic dependent launch of the library's kernels (every kernel opens with griddepcontrol.launch_dependents +
 * griddepcontrol.wait, so the launch latency of kernel N+1 hides under kernel N; stream order is unchanged):
 * 1 (default, env SCN_B200_PDL): every kernel; 2: all but the persistent tensor-core GEMMs; 0: plain launches */
int scn_set_pdl(int mode);
/* CUDA events as opaque handles (timing disabled) for the marks above */
int scn_event_create(void **event);
int scn_event_destroy(void *event);
int scn_event_record(void *event, void *stream);
int scn_stream_wait_event(void *stream, void *event);

/* ---- elementwise helper used by the data-parallel gradient bucket ------------------- */
/* y[i] *= alpha over n fp32 values (scale the all-reduced gradient bucket by 1/world) */
int scn_scale_inplace(float *y, float alpha, int64_t n, void *stream);

/* ---- instrumentation --------------------------------------------------------------- */
/* number of kernels this library has launched since load (bench.py "gpu_launches") */
int64_t scn_launch_count(void);
/* per-kernel-class timing for the roofline report: when enabled every region of class `cls`
 * (0 conv gather-GEMM fwd/dX, 1 weight gradient, 2 batch-norm, 3 hash/rulebook build,
 * 4 input/output/sparse-to-dense) is bracketed by CUDA events on its stream.  scn_prof_read
 * synchronises, returns out = [regions, total ms, algorithmic bytes, flops] and clears them. */
int scn_prof_enable(int on);
int scn_prof_read(int cls, double out[4]);
/* tile-book statistics of a cached rulebook: stats[0]=pairs, [1]=tile entries*128 (rows the
 * tensor/FFMA tiles actually multiply), [2]=tiles.  kind: 0 submanifold, 1 conv, 2 deconv,
 * 3 conv-dX (fine<-coarse), 4 deconv-dX */
int scn_rulebook_stats(scn_metadata_t *m, int kind, const int64_t *in_spatial_size,
                       const int64_t *filter_size, const int64_t *filter_stride,
                       int64_t stats[3]);
/* test / tuning knob: cap the grid of the persistent gather-GEMM (0 = one CTA per SM), so that small
 * inputs exercise the many-work-items-per-CTA paths */
int scn_set_gemm_grid_limit(int max_ctas);
/* 0: tiles in natural row order, 1 (default): rows grouped by neighbour mask */
int scn_set_tile_grouping(int enabled);

/* ---- ROIAlignRotated3D sampled from the sparse map -----------------------
 * Replaces the pair  sparse_3d_to_dense_2d(input_s3d)  (SparseConvNet/sparseconvnet/tools_3d_2d.py:7-26: a dense
 * [B,C,X,Y,Z] tensor cropped to max active coordinate + 1)  +  _C.roi_align_rotated_3d_forward / _backward
 * (maskrcnn_benchmark/layers/roi_align_rotated_3d.py:14-57 ->
 * maskrcnn_benchmark/csrc/ROIAlignRotated3D.h, csrc/cuda/ROIAlignRotated3D_cuda.cu:357-454): the trilinear
 * corners are looked up in the scale's hash grid, inactive corners contribute 0 (= the zero-filled dense tensor).
 * feats [nActive(ss), n_planes] device; rois [n_rois, 8] device float (batch, center_w, center_h, center_z,
 * width, height, zsize, theta in degrees - same meaning and axis convention as the reference: the dense H axis is
 * the sparse x axis, W the sparse y axis); pooled = {pooled_height, pooled_width, pooled_zsize};
 * out [n_rois, n_planes, pooled_height, pooled_width, pooled_zsize] device.
 * backward: d_feats [nActive, n_planes] is zero-filled, then receives the gradient of the ACTIVE sites (what
 * SparseToDense's backward keeps of the reference's dense gradient). */
int scn_roi_align_rotated_3d_forward(scn_metadata_t *m, const int64_t *spatial_size, const float *feats,
                                     int64_t n_planes, const float *rois, int64_t n_rois, float spatial_scale,
                                     const int64_t *pooled, int sampling_ratio, float *out, void *stream);
int scn_roi_align_rotated_3d_backward(scn_metadata_t *m, const int64_t *spatial_size, const float *d_out,
                                      int64_t n_planes, const float *rois, int64_t n_rois, float spatial_scale,
                                      const int64_t *pooled, int sampling_ratio, float *d_feats, void *stream);

/* ---- RPN: anchors and head on the device ----------------------------------
 * scn_grid_anchors replaces AnchorGenerator.grid_anchors + examples_bidx_2_sizes
 * (maskrcnn_benchmark/modeling/rpn/anchor_generator_sparse3d.py:88-104, 137-146, 174-185), which copy every
 * level's get_spatial_locations() to the host each step: anchors_out DEVICE float [nActive(ss) * n_anchors, 7] =
 * ((x,y,z) / voxel_scale * stride, 0,0,0,0) + base_anchors[a] in the reference's flatten order [site, anchor, 7];
 * scope_out (optional) DEVICE int64 [batch_size, 2] = row range of every sample times n_anchors.
 * base_anchors DEVICE float [n_anchors, 7]; stride HOST float[3]. */
int scn_grid_anchors(scn_metadata_t *m, const int64_t *spatial_size, const float *base_anchors, int64_t n_anchors,
                     float voxel_scale, const float *stride, float *anchors_out, int64_t *scope_out,
                     int64_t batch_size, void *stream);
/* RPNHead.forward (maskrcnn_benchmark/modeling/rpn/rpn_sparse3d.py:97-131) on the row-major [n, C] features of one
 * level: hidden = relu(x Wc^T + bc), logits [n, n_cls] = hidden Wl^T + bl, reg [n, n_box] = hidden Wr^T + br
 * (n_cls = A * S, n_box = 7 * A * S; row-major [n, A*S] IS the reference's [1, n, A, S] after its permute + reshape).
 * Weights in nn.Conv2d layout [Cout, Cin(,1,1)], consumed in place.  `hidden` [n, C] is an output the caller keeps
 * for the backward pass; d_hidden [n, C] is caller-provided scratch; d_x may be NULL. */
int scn_rpn_head_forward(const float *x, int64_t n, int64_t n_planes, const float *w_conv, const float *b_conv,
                         const float *w_cls, const float *b_cls, int64_t n_cls, const float *w_box,
                         const float *b_box, int64_t n_box, float *hidden, float *logits, float *reg, int precision,
                         void *stream);
int scn_rpn_head_backward(const float *x, const float *hidden, int64_t n, int64_t n_planes, const float *w_conv,
                          const float *w_cls, int64_t n_cls, const float *w_box, int64_t n_box,
                          const float *d_logits, const float *d_reg, float *d_hidden, float *d_x, float *dw_conv,
                          float *db_conv, float *dw_cls, float *db_cls, float *dw_box, float *db_box, int precision,
                          void *stream);

/* ---- rotated box IoU / rotated NMS on the device -------------------------
 * scn_rotate_iou replaces second/core/non_max_suppression/nms_gpu.py:667-703 rotate_iou_gpu_eval (numba-CUDA,
 * host arrays in / out): boxes DEVICE float [n_boxes,5], query DEVICE float [n_query,5], rows (x, y, size_x,
 * size_y, yaw); iou_out DEVICE float [n_boxes, n_query]; criterion as in devRotateIoUEval (:552-623: -1 IoU,
 * 0 inter / query area, 1 inter / box area, 2 thin-box variant, 3-6 its experimental distance scores, else the
 * intersection area).
 * scn_rotate_nms replaces the body of rotate_nms_3d (second/pytorch/core/box_torch_ops.py:557-582 ->
 * second/core/non_max_suppression/nms_cpu.py:32-44 -> spconv.utils.rotate_non_max_suppression_cpu): top
 * pre_max_size boxes by score (<= 0: all), greedy suppression in score order where IoU >= iou_threshold, at most
 * post_max_size kept (<= 0: no limit).  boxes DEVICE float [n,5] BEV rows as above (the caller passes
 * bbox3d[:, [0,1,3,4,6]] as the reference does), scores DEVICE float [n]; keep_out DEVICE int64 [min(n, post)]
 * indices into the input in descending score order; *n_keep read back to the host (one sync). */
int scn_rotate_iou(const float *boxes, int64_t n_boxes, const float *query, int64_t n_query, int criterion,
                   float *iou_out, void *stream);
int scn_rotate_nms(const float *boxes, const float *scores, int64_t n, float iou_threshold, int64_t pre_max_size,
                   int64_t post_max_size, int64_t *keep_out, int64_t *n_keep, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SCN_B200_H */
