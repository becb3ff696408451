// trainer_capi.cpp — C ABI around one training step of the reference loop
// (recipes/slimIPL/src/Train.cpp:1454-1803), written against fl_compat exactly as Train.cpp is written
// against flashlight:
//   input -> ntwrk->forward -> criterion->forward -> zeroGrad -> loss.backward() -> reducer (NCCL all-reduce
//   of every net + criterion gradient) -> grads / (totalBatch) -> clipGradNorm(net U crit) -> critopt/netopt step
// B200-first differences, all behind the same call sequence: parameters, gradients and momentum live in flat
// arenas (one NCCL call, one norm kernel, one fused scale+clip+SGD kernel instead of per-array JIT kernels).
// The Python harness (bench.py, tests) drives it with device pointers; no arithmetic happens on the host.
#include <cuda_runtime.h>

#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "fl_compat/fl_compat.h"
#include "w2l_b200.h"

namespace w2l {
int fail(int code, const std::string& msg);
void check(int rc);
}  // namespace w2l

using namespace fl;
using namespace fl::pkg::speech;

namespace {
struct Trainer {
  std::shared_ptr<fl::Module> net;
  std::unique_ptr<fl::OverlappedArenaReducer> reducer;  // created at the first distributed step
  std::shared_ptr<SequenceCriterion> crit;
  ParameterArena netArena, critArena;
  af::array sqnorm;
  af::array guard;  // int32[2]: [0] this step had a non-finite loss / gradient (update skipped), [1] count of such steps
  int precision = W2L_PRECISION_TF32;
  float lr, lrcrit, momentum, maxgradnorm;
  int nFeat, nLabel;
  bool isCtc;
};

struct PrecisionScope {  // the trainer's precision for the duration of one call; the thread's own setting is restored
  int saved;
  explicit PrecisionScope(int p) : saved(w2l_get_precision()) { w2l_set_precision(p); }
  ~PrecisionScope() { w2l_set_precision(saved); }
};

template <typename F>
int guarded(F&& f) {
  try {
    f();
    return W2L_OK;
  } catch (const std::invalid_argument& e) {
    return w2l::fail(W2L_ERR_INVALID_ARGUMENT, e.what());
  } catch (const std::exception& e) {
    return w2l::fail(W2L_ERR_CUDA, e.what());
  }
}
}  // namespace

extern "C" {

W2L_API void* w2l_trainer_create(void* stream, const char* arch_text, int n_feat, int n_label, const char* criterion, int scale_mode,
                                 float transdiag, float lr, float lrcrit, float momentum, float maxgradnorm) {
  Trainer* t = nullptr;
  const int rc = guarded([&] {
    w2l::setCurrentStream(stream);
    auto tr = std::make_unique<Trainer>();
    // --arch is either the text of an .arch file or the path of a plugin exporting createModule (Train.cpp:390-395)
    const std::string arch = arch_text;
    if (arch.size() > 3 && arch.compare(arch.size() - 3, 3, ".so") == 0 && arch.find('\n') == std::string::npos)
      tr->net = fl::pkg::runtime::ModulePlugin(arch).arch(n_feat, n_label);
    else
      tr->net = fl::pkg::runtime::buildSequentialModule(arch, n_feat, n_label);
    const auto mode = static_cast<CriterionScaleMode>(scale_mode);
    const std::string c = criterion;
    if (c == "ctc") {
      tr->crit = std::make_shared<CTCLoss>(mode);
      tr->isCtc = true;
    } else if (c == "asg") {
      tr->crit = std::make_shared<ASGLoss>(n_label, mode, transdiag);
      tr->isCtc = false;
    } else if (c == "linseg") {
      // the --linseg warm start (Train.cpp:589-617, :1867-1883): LinSegCriterion sharing the ASG transition matrix
      auto asg = std::make_shared<ASGLoss>(n_label, mode, transdiag);
      auto lin = std::make_shared<LinSegCriterion>(n_label, mode);
      lin->setParams(asg->param(0), 0);
      tr->crit = lin;
      tr->isCtc = false;
    } else {
      throw std::invalid_argument("criterion must be 'ctc', 'asg' or 'linseg'");
    }
    tr->netArena = flattenParameters({tr->net});
    if (!tr->crit->params().empty()) tr->critArena = flattenParameters({tr->crit});
    tr->sqnorm = af::array::zeros(af::dim4(1), w2l::DType::f64);
    tr->guard = af::array::zeros(af::dim4(2), w2l::DType::i32);
    tr->precision = w2l_get_precision();  // the creating thread's setting; w2l_trainer_set_precision changes it
    tr->lr = lr;
    tr->lrcrit = lrcrit;
    tr->momentum = momentum;
    tr->maxgradnorm = maxgradnorm;
    tr->nFeat = n_feat;
    tr->nLabel = n_label;
    t = tr.release();
  });
  return rc == W2L_OK ? t : nullptr;
}

W2L_API void w2l_trainer_destroy(void* h) { delete static_cast<Trainer*>(h); }

W2L_API long long w2l_trainer_num_params(void* h, int which /*0 net, 1 criterion*/) {
  auto* t = static_cast<Trainer*>(h);
  return which == 0 ? t->netArena.elements : t->critArena.elements;
}

// copy the flat arena (what: 0 values, 1 gradients) to a device buffer of num_params floats
W2L_API int w2l_trainer_get_flat(void* h, void* stream, int which, int what, float* out_dev) {
  return guarded([&] {
    w2l::setCurrentStream(stream);
    auto* t = static_cast<Trainer*>(h);
    const ParameterArena& a = which == 0 ? t->netArena : t->critArena;
    if (a.elements == 0) return;
    af::array dst = af::array::wrap(out_dev, af::dim4(a.elements));
    dst.copyFrom(what == 0 ? a.values : a.grads);
  });
}
W2L_API int w2l_trainer_set_flat(void* h, void* stream, int which, const float* in_dev) {
  return guarded([&] {
    w2l::setCurrentStream(stream);
    auto* t = static_cast<Trainer*>(h);
    const ParameterArena& a = which == 0 ? t->netArena : t->critArena;
    if (a.elements == 0) return;
    a.values.copyFrom(af::array::wrap(const_cast<float*>(in_dev), af::dim4(a.elements)));
  });
}
// layout of the arena: for parameter i, elements and 4 dims; returns the parameter count
W2L_API int w2l_trainer_param_layout(void* h, int which, int max_params, long long* elements, long long* dims4) {
  auto* t = static_cast<Trainer*>(h);
  auto ps = which == 0 ? t->net->params() : t->crit->params();
  for (int i = 0; i < (int)ps.size() && i < max_params; ++i) {
    elements[i] = ps[i].elements();
    for (int d = 0; d < 4; ++d) dims4[4 * i + d] = ps[i].dims(d);
  }
  return (int)ps.size();
}

// One step.  features: device [T,F,1,B] (ArrayFire layout, T fastest), target: device [L,B] int32 (-1 padded),
// loss_out: device [B].  train != 0 runs backward + all-reduce + clip + SGD.  total_batch = sum of B over ranks.
W2L_API int w2l_trainer_step(void* h, void* stream, int B, int T, const float* features, int L, const int32_t* target,
                             float* loss_out, int train, float total_batch) {
  return guarded([&] {
    w2l::setCurrentStream(stream);
    auto* t = static_cast<Trainer*>(h);
    PrecisionScope scope(t->precision);
    if (train) {
      t->net->train();
      t->crit->train();
    } else {
      t->net->eval();
      t->crit->eval();
    }
    // [fwd]  Train.cpp:1454-1470
    Variable input = fl::input(af::array::wrap(const_cast<float*>(features), af::dim4(T, t->nFeat, 1, B)));
    Variable output = t->net->forward(std::vector<Variable>{input}).front();
    // [crit] Train.cpp:1675
    Variable tgt = fl::noGrad(af::array::wrap(const_cast<int32_t*>(target), af::dim4(L, B), w2l::DType::i32));
    Variable loss = t->crit->forward({output, tgt}).front();
    if (loss_out) af::array::wrap(loss_out, af::dim4(B)).copyFrom(loss.array());
    if (!train) return;
    // [bwd]  zeroGrad; loss.backward()   Train.cpp:1718-1720
    t->netArena.grads.zero();
    for (auto& p : t->net->params()) p.zeroGrad(false);
    if (t->critArena.elements) {
      t->critArena.grads.zero();
      for (auto& p : t->crit->params()) p.zeroGrad(false);
    }
    // reducer: Train.cpp:1721-1735 adds every gradient after backward; here the network's gradient arena is reduced in
    // buckets WHILE backward runs (the callback pattern of cpc/Train.cpp:972-976), on a separate stream
    if (fl::isDistributedInit() && !t->reducer) t->reducer = std::make_unique<fl::OverlappedArenaReducer>(t->net->params(), t->netArena.grads);
    if (t->reducer) t->reducer->arm();
    loss.backward();
    if (t->reducer) t->reducer->finalize();
    if (fl::isDistributedInit() && t->critArena.elements) fl::allReduce(t->critArena.grads);
    // [opt]  grads /= totalBatch (Train.cpp:1752,1783), clipGradNorm(net U crit) (:1791-1798), step (:1801-1802).
    // The reference's numerical guards — LOG(FATAL) on a NaN / Inf loss (:1686-1698), skip-and-retry on non-finite
    // gradients under mixed precision (:1753-1771) — run on the device: the squared gradient norm is computed anyway, a
    // one-CTA kernel turns "loss or norm not finite" into a flag, and both SGD kernels return early when it is set.
    // w2l_trainer_status reads the count of skipped steps whenever the caller wants it (no sync inside the step).
    const float gscale = 1.0f / total_batch;
    t->sqnorm.zero();
    w2l::check(w2l_sq_norm_accumulate(stream, t->netArena.elements, t->netArena.grads.f32(), t->sqnorm.f64()));
    if (t->critArena.elements)
      w2l::check(w2l_sq_norm_accumulate(stream, t->critArena.elements, t->critArena.grads.f32(), t->sqnorm.f64()));
    w2l::check(w2l_finite_guard(stream, B, loss.array().f32(), t->sqnorm.f64(), t->guard.i32()));
    const double* sq = t->maxgradnorm > 0 ? t->sqnorm.f64() : nullptr;
    if (t->critArena.elements)
      w2l::check(w2l_sgd_step_ex(stream, t->critArena.elements, t->critArena.values.f32(), t->critArena.grads.f32(), t->critArena.velocity.f32(),
                                 t->lrcrit, 0.f, 0.f, gscale, t->maxgradnorm, sq, 0, t->guard.i32()));
    w2l::check(w2l_sgd_step_ex(stream, t->netArena.elements, t->netArena.values.f32(), t->netArena.grads.f32(), t->netArena.velocity.f32(), t->lr,
                               t->momentum, 0.f, gscale, t->maxgradnorm, sq, 0, t->guard.i32()));
  });
}

W2L_API int w2l_trainer_set_precision(void* h, int precision) {
  if (precision != W2L_PRECISION_TF32 && precision != W2L_PRECISION_F32 && precision != W2L_PRECISION_BF16)
    return w2l::fail(W2L_ERR_INVALID_ARGUMENT, "trainer_set_precision: unknown precision");
  static_cast<Trainer*>(h)->precision = precision;
  return W2L_OK;
}
// number of steps whose update was skipped because the loss or a gradient was NaN / Inf (synchronises the stream)
W2L_API int w2l_trainer_status(void* h, void* stream, long long* skipped_steps) {
  return guarded([&] {
    w2l::setCurrentStream(stream);
    auto* t = static_cast<Trainer*>(h);
    const std::vector<int32_t> g = t->guard.host<int32_t>();
    if (skipped_steps) *skipped_steps = g[1];
  });
}

// network forward only (eval mode): emissions_out device [N,T',B]; returns T' through t_out
W2L_API int w2l_trainer_forward(void* h, void* stream, int B, int T, const float* features, float* emissions_out, long long capacity,
                                int* t_out) {
  return guarded([&] {
    w2l::setCurrentStream(stream);
    auto* t = static_cast<Trainer*>(h);
    PrecisionScope scope(t->precision);
    t->net->eval();
    Variable out = t->net->forward(std::vector<Variable>{fl::input(af::array::wrap(const_cast<float*>(features), af::dim4(T, t->nFeat, 1, B)))}).front();
    if (out.elements() > capacity) throw std::invalid_argument("trainer_forward: output buffer too small");
    af::array::wrap(emissions_out, out.dims()).copyFrom(out.array());
    if (t_out) *t_out = (int)(out.elements() / ((long long)B * t->nLabel));  // frames of nLabel values per sample
  });
}

W2L_API int w2l_nccl_unique_id(void* out128) {
  return guarded([&] { fl::pkg::runtime::createUniqueId(out128); });
}
W2L_API int w2l_init_distributed(int rank, int world, const void* id128) {
  return guarded([&] { fl::pkg::runtime::initDistributed(rank, world, id128); });
}
W2L_API int w2l_trainer_sync_parameters(void* h, void* stream) {  // fl::allReduceParameters, Train.cpp:1078-1079
  return guarded([&] {
    w2l::setCurrentStream(stream);
    auto* t = static_cast<Trainer*>(h);
    if (!fl::isDistributedInit()) return;
    fl::allReduce(t->netArena.values, 1.0 / fl::getWorldSize());
    if (t->critArena.elements) fl::allReduce(t->critArena.values, 1.0 / fl::getWorldSize());
  });
}
W2L_API const char* w2l_trainer_describe(void* h) {
  static thread_local std::string s;
  auto* t = static_cast<Trainer*>(h);
  s = t->net->prettyString() + "\n" + t->crit->prettyString();
  return s.c_str();
}
}
