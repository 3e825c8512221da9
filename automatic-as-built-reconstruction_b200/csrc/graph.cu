// graph.cu - layer-graph executor: the backbone's ~130 layer ops per direction run from ONE foreign call
// each (include/scn_b200.h "layer-graph executor").  It calls exactly the per-op entry points the Python
// layer files call (same kernels, same results); what it removes is the Python / autograd enqueue work
// between them, which at batch 1 bounds the step (reference dataflow: sparseconvnet/fpn_net.py:168-265,
// sequential.py:15-17, tables.py:28-56).
#include "common.cuh"
#include "conv.cuh"
#include "../../include/scn_b200.h"
#include <vector>

namespace scn {

__global__ void k_add(const float *a, const float *b, float *out, long long n4, long long n) {
  pdl_sync();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long st = (long long)gridDim.x * blockDim.x;
  for (long long q = i; q < n4; q += st) {
    const float4 x = reinterpret_cast<const float4 *>(a)[q], y = reinterpret_cast<const float4 *>(b)[q];
    reinterpret_cast<float4 *>(out)[q] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
  }
  for (long long q = n4 * 4 + i; q < n; q += st) out[q] = a[q] + b[q];
}

// out = a + b over n floats (out may alias a or b)
static int add_into(const float *a, const float *b, float *out, long long n, cudaStream_t s) {
  if (n <= 0) return 0;
  const bool al = (((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) == 0;
  const long long n4 = al ? n / 4 : 0;
  long long blocks = ((al ? n4 : n) + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  SCN_LAUNCH(k_add, (int)blocks, 256, 0, s, a, b, out, n4, n);
  SCN_LAUNCHED();
  return 0;
}

static bool g_graph_defer_dw = true;   // scn_set_graph_overlap (the roofline pass of bench.py turns it off)

static const int64_t *tag_of(const int64_t *tags, int p) {
  return (tags && p >= 0 && tags[2 * p] != 0) ? tags + 2 * p : nullptr;
}

}  // namespace scn

using namespace scn;

extern "C" {

int scn_graph_forward(scn_metadata_t *m, const scn_graph_op_t *ops, int32_t n_ops, float *const *values,
                      const int64_t *rows, float *const *params, const int64_t *param_tags, float *bn_save,
                      int train, int precision, void *stream, double *macs) {
  SCN_CHECK(m && ops && values && rows && params, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  double total = 0;
  struct BatchEnd { ~BatchEnd() { prepack_weights_end(); } } batch_end;     // (also on the error paths)
  {   // operand images of every convolution weight in one launch (the per-op calls below then find them fresh)
    std::vector<const int64_t *> tags;
    std::vector<const float *> ws;
    std::vector<int> Ks, cins, couts;
    for (int i = 0; i < n_ops; ++i) {
      const scn_graph_op_t &o = ops[i];
      if (o.kind < 1 || o.kind > 3 || o.p0 < 0) continue;
      tags.push_back(tag_of(param_tags, o.p0));
      ws.push_back(params[o.p0]);
      Ks.push_back((int)(o.filter[0] * o.filter[1] * o.filter[2]));
      cins.push_back(o.n_in_planes);
      couts.push_back(o.n_out_planes);
    }
    SCN_TRY(prepack_weights_batch((int)ws.size(), tags.data(), ws.data(), Ks.data(), cins.data(), couts.data(), precision, s));
  }
  for (int i = 0; i < n_ops; ++i) {
    const scn_graph_op_t &o = ops[i];
    double mac = 0;
    const float *w = o.p0 >= 0 ? params[o.p0] : nullptr;
    const float *b = o.p1 >= 0 ? params[o.p1] : nullptr;
    switch (o.kind) {
      case 1:
        SCN_TRY(scn_submanifold_conv_forward(m, o.in_ss, o.filter, values[o.in0], values[o.out], w, b, o.n_in_planes,
                                             o.n_out_planes, precision, stream, &mac, tag_of(param_tags, o.p0)));
        break;
      case 2:
        SCN_TRY(scn_conv_forward(m, o.in_ss, o.out_ss, o.filter, o.stride, values[o.in0], values[o.out], w, b,
                                 o.n_in_planes, o.n_out_planes, precision, stream, &mac, tag_of(param_tags, o.p0)));
        break;
      case 3:
        SCN_TRY(scn_deconv_forward(m, o.in_ss, o.out_ss, o.filter, o.stride, values[o.in0], values[o.out], w, b,
                                   o.n_in_planes, o.n_out_planes, precision, stream, &mac, tag_of(param_tags, o.p0)));
        break;
      case 4:
        SCN_CHECK(bn_save && o.p2 >= 0 && o.p3 >= 0, "graph op %d: BatchNorm needs running statistics and a save area", i);
        SCN_TRY(scn_batchnorm_forward(values[o.in0], values[o.out], bn_save + o.save_off,
                                      bn_save + o.save_off + o.n_out_planes, params[o.p2], params[o.p3], w, b, o.eps,
                                      o.momentum, train, o.leakiness, rows[o.in0], o.n_out_planes, stream));
        break;
      case 5:
        SCN_CHECK(rows[o.in0] == rows[o.in1] && rows[o.in0] == rows[o.out], "graph op %d: add of unequal row counts", i);
        SCN_TRY(add_into(values[o.in0], values[o.in1], values[o.out], rows[o.out] * (long long)o.n_out_planes, s));
        break;
      default:
        SCN_CHECK(false, "graph op %d: unknown kind %d", i, o.kind);
    }
    total += mac;
  }
  if (macs) *macs = total;
  return 0;
}

int scn_graph_backward(scn_metadata_t *m, const scn_graph_op_t *ops, int32_t n_ops, int32_t n_values,
                       float *const *values, const int64_t *rows, float *const *params,
                       const int64_t *param_tags, const float *bn_save, float *const *grads,
                       const float *const *out_grads, float *const *param_grads, uint8_t *param_written,
                       float *scratch, int64_t scratch_floats, int precision, void *stream) {
  return scn_graph_backward_marked(m, ops, n_ops, n_values, values, rows, params, param_tags, bn_save, grads, out_grads,
                                   param_grads, param_written, scratch, scratch_floats, precision, stream, 0, nullptr,
                                   nullptr);
}

int scn_set_graph_overlap(int mode) {
  g_graph_defer_dw = mode >= 2;
  g_dw_companion = mode >= 1;
  return 0;
}

int scn_event_create(void **event) {
  SCN_CHECK(event, "null argument");
  cudaEvent_t e;
  SCN_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  *event = (void *)e;
  return 0;
}
int scn_event_destroy(void *event) {
  if (event) SCN_CUDA(cudaEventDestroy((cudaEvent_t)event));
  return 0;
}
int scn_event_record(void *event, void *stream) {
  SCN_CHECK(event, "null argument");
  SCN_CUDA(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
  return 0;
}
int scn_stream_wait_event(void *stream, void *event) {
  SCN_CHECK(event, "null argument");
  SCN_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0));
  return 0;
}

int scn_graph_backward_marked(scn_metadata_t *m, const scn_graph_op_t *ops, int32_t n_ops, int32_t n_values,
                              float *const *values, const int64_t *rows, float *const *params,
                              const int64_t *param_tags, const float *bn_save, float *const *grads,
                              const float *const *out_grads, float *const *param_grads, uint8_t *param_written,
                              float *scratch, int64_t scratch_floats, int precision, void *stream, int32_t n_marks,
                              const int32_t *mark_ops, void *const *mark_events) {
  SCN_CHECK(m && ops && values && rows && params && grads && out_grads && param_grads && param_written, "null argument");
  SCN_CHECK(n_marks == 0 || (mark_ops && mark_events), "null mark arrays");
  cudaStream_t s = (cudaStream_t)stream;
  // marks: event j is recorded on the stream once the reverse sweep has passed op mark_ops[j] (whether that op ran
  // or was skipped as a dead branch): every parameter gradient of the ops >= mark_ops[j] is then final
  int next_mark = 0;
  auto fire_marks = [&](int op_done) -> int {
    while (next_mark < n_marks && mark_ops[next_mark] >= op_done) {
      SCN_TRY(dw_join_pending(s));                     // the weight gradients queued on the companion stream so far
      SCN_CUDA(cudaEventRecord((cudaEvent_t)mark_events[next_mark], s));
      ++next_mark;
    }
    return 0;
  };
  // weight gradients run on the companion stream WITHOUT a join per layer: they only have to be complete at the
  // marks and when the sweep returns (the join also runs on the error paths)
  struct DeferJoin {
    cudaStream_t s;
    explicit DeferJoin(cudaStream_t st) : s(st) { g_defer_dw_join = g_graph_defer_dw; }
    ~DeferJoin() { g_defer_dw_join = false; dw_join_pending(s); }
  } defer_join(s);
  for (int j = 1; j < n_marks; ++j) SCN_CHECK(mark_ops[j] <= mark_ops[j - 1], "marks must be in descending op order");
  // current gradient of every value: none yet / someone else's finished buffer (read only) / its own buffer
  enum { NONE = 0, ALIAS = 1, OWN = 2 };
  struct G { int st; const float *p; };
  std::vector<G> g((size_t)n_values, G{NONE, nullptr});
  for (int v = 0; v < n_values; ++v)
    if (out_grads[v]) g[v] = G{ALIAS, out_grads[v]};
  std::vector<int> width((size_t)n_values, 0);
  for (int i = 0; i < n_ops; ++i) {
    width[ops[i].in0] = ops[i].n_in_planes;
    if (ops[i].kind == 5) width[ops[i].in1] = ops[i].n_in_planes;
    width[ops[i].out] = ops[i].n_out_planes;
  }
  // a gradient that passes through unchanged (add): alias it, or add it to what is already there
  auto pass = [&](int v, const float *p) -> int {
    if (!grads[v]) return 0;
    if (g[v].st == NONE) { g[v] = G{ALIAS, p}; return 0; }
    SCN_TRY(add_into(g[v].p, p, grads[v], rows[v] * (long long)width[v], s));
    g[v] = G{OWN, grads[v]};
    return 0;
  };
  for (int i = n_ops - 1; i >= 0; --i) {
    SCN_TRY(fire_marks(i + 1));                        // everything above op i is finished
    const scn_graph_op_t &o = ops[i];
    if (g[o.out].st == NONE) continue;                 // dead branch: no gradient reaches this op
    const float *dY = g[o.out].p;
    const int v = o.in0;
    if (o.kind == 5) {
      SCN_TRY(pass(o.in0, dY));
      SCN_TRY(pass(o.in1, dY));
      continue;
    }
    const float *w = o.p0 >= 0 ? params[o.p0] : nullptr;
    float *dw = o.p0 >= 0 ? param_grads[o.p0] : nullptr;
    float *db = o.p1 >= 0 ? param_grads[o.p1] : nullptr;
    if (o.kind == 4) {
      SCN_CHECK(grads[v], "graph op %d: BatchNorm input needs a gradient buffer", i);
      // BN gradient + whatever already arrived for the same value (skip connection), in one pass
      SCN_TRY(scn_batchnorm_backward_fused(values[v], grads[v], values[o.out], dY, bn_save + o.save_off,
                                           bn_save + o.save_off + o.n_out_planes, w, o.p1 >= 0 ? params[o.p1] : nullptr,
                                           1, dw, db, o.leakiness, rows[v], o.n_out_planes,
                                           g[v].st == NONE ? nullptr : g[v].p, stream));
      g[v] = G{OWN, grads[v]};
    } else {
      float *d_in = nullptr;
      bool accumulate = false;
      if (grads[v]) {
        if (g[v].st == NONE) d_in = grads[v];
        else {
          SCN_CHECK(scratch && scratch_floats >= rows[v] * (long long)width[v], "graph op %d: scratch too small", i);
          d_in = scratch;
          accumulate = true;
        }
      }
      const int64_t *tag = tag_of(param_tags, o.p0);
      if (o.kind == 1)
        SCN_TRY(scn_submanifold_conv_backward(m, o.in_ss, o.filter, values[v], d_in, dY, w, dw, db, o.n_in_planes,
                                              o.n_out_planes, precision, stream, tag));
      else if (o.kind == 2)
        SCN_TRY(scn_conv_backward(m, o.in_ss, o.out_ss, o.filter, o.stride, values[v], d_in, dY, w, dw, db,
                                  o.n_in_planes, o.n_out_planes, precision, stream, tag));
      else if (o.kind == 3)
        SCN_TRY(scn_deconv_backward(m, o.in_ss, o.out_ss, o.filter, o.stride, values[v], d_in, dY, w, dw, db,
                                    o.n_in_planes, o.n_out_planes, precision, stream, tag));
      else
        SCN_CHECK(false, "graph op %d: unknown kind %d", i, o.kind);
      if (accumulate) SCN_TRY(add_into(g[v].p, scratch, grads[v], rows[v] * (long long)width[v], s));
      if (grads[v]) g[v] = G{OWN, grads[v]};
    }
    if (o.p0 >= 0 && dw) param_written[o.p0] = 1;
    if (o.p1 >= 0 && db) param_written[o.p1] = 1;
  }
  SCN_TRY(fire_marks(0));
  SCN_TRY(dw_join_pending(s));
  return 0;
}

}  // extern "C"
