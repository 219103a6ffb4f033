// oracle/ref_model_driver.cpp -- TEST INFRASTRUCTURE ONLY.
// Drives the UNMODIFIED reference Model (model.cpp compiled from /root/reference by oracle/Makefile) so that the
// restatements in oracle/scene_edit.py can be pinned against the reference itself:
//   torch.ops.opensplat_ref_model.after_train   -> Model::afterTrain  (model.cpp:311-500), incl. the Adam-state surgery
//   torch.ops.opensplat_ref_model.save          -> Model::save        (model.cpp:496-594; .ply or .splat by extension)
// Nothing of the reference is copied: a Model is constructed through its own constructor (with a stub for the
// nanoflann-based PointsTensor::scales(), whose result we overwrite anyway), its public tensors are replaced by the
// caller's, its optimizers are re-created by its own setupOptimizers(), and the Adam moments are injected as the
// AdamParamState entries that Model::addToOptimizer / removeFromOptimizer read.
#include <torch/torch.h>
#include <torch/library.h>

#include "model.hpp"

// ---- symbols of reference translation units we do not link (never reached by the calls above) ----
torch::Tensor PointsTensor::scales() { return torch::ones({tensor.size(0), 1}, torch::kFloat32); }
PointsTensor::~PointsTensor() {}

namespace {

using torch::Tensor;

void inject_state(torch::optim::Adam *opt, const Tensor &m, const Tensor &v) {
    Tensor param = opt->param_groups()[0].params()[0];
    auto st = std::make_unique<torch::optim::AdamParamState>();
    st->step(1);
    st->exp_avg(m.clone());
    st->exp_avg_sq(v.clone());
    opt->state()[param.unsafeGetTensorImpl()] = std::move(st);
}

std::pair<Tensor, Tensor> read_state(torch::optim::Adam *opt) {
    Tensor param = opt->param_groups()[0].params()[0];
    auto &st = static_cast<torch::optim::AdamParamState &>(*opt->state()[param.unsafeGetTensorImpl()]);
    return {st.exp_avg().clone(), st.exp_avg_sq().clone()};
}

std::unique_ptr<Model> make_model(const std::vector<Tensor> &params, int64_t numCameras, int64_t refineEvery,
                                  int64_t warmupLength, int64_t resetAlphaEvery, double densifyGradThresh,
                                  double densifySizeThresh, int64_t stopScreenSizeAt, double splitScreenSize,
                                  int64_t maxSteps, bool keepCrs, double scale, const Tensor &translation) {
    const int64_t n = params[0].size(0);
    const int64_t restBases = params[4].size(1);
    int shDegree = 0;
    while ((shDegree + 1) * (shDegree + 1) < restBases + 1) ++shDegree;
    InputData in;
    in.scale = (float)scale;
    in.translation = translation.clone();
    in.points.xyz = torch::zeros({n, 3}, torch::kFloat32);
    in.points.rgb = torch::zeros({n, 3}, torch::kUInt8);
    auto m = std::make_unique<Model>(in, (int)numCameras, /*numDownscales*/ 0, /*resolutionSchedule*/ 3000, shDegree,
                                     /*shDegreeInterval*/ 1000, (int)refineEvery, (int)warmupLength,
                                     (int)resetAlphaEvery, (float)densifyGradThresh, (float)densifySizeThresh,
                                     (int)stopScreenSizeAt, (float)splitScreenSize, (int)maxSteps, keepCrs,
                                     torch::Device(torch::kCPU));
    m->means = params[0].clone().requires_grad_();
    m->scales = params[1].clone().requires_grad_();
    m->quats = params[2].clone().requires_grad_();
    m->featuresDc = params[3].clone().requires_grad_();
    m->featuresRest = params[4].clone().requires_grad_();
    m->opacities = params[5].clone().requires_grad_();
    m->releaseOptimizers();
    m->setupOptimizers();
    return m;
}

// params = {means, scales, quats, featuresDc, featuresRest, opacities}; adam_m / adam_v in the same order;
// stats = {} (first step after a clear) or {xysGradNorm, visCounts, max2DSize}.
// Returns params(6) + adam_m(6) + adam_v(6) + stats(3, empty tensors when cleared) after Model::afterTrain(step).
std::vector<Tensor> after_train(std::vector<Tensor> params, std::vector<Tensor> adam_m, std::vector<Tensor> adam_v,
                                std::vector<Tensor> stats, Tensor xys_grad, Tensor radii, int64_t height,
                                int64_t width, int64_t step, int64_t seed, int64_t numCameras, int64_t refineEvery,
                                int64_t warmupLength, int64_t resetAlphaEvery, double densifyGradThresh,
                                double densifySizeThresh, int64_t stopScreenSizeAt, double splitScreenSize,
                                int64_t maxSteps) {
    auto m = make_model(params, numCameras, refineEvery, warmupLength, resetAlphaEvery, densifyGradThresh,
                        densifySizeThresh, stopScreenSizeAt, splitScreenSize, maxSteps, false, 1.0,
                        torch::zeros({3}, torch::kFloat32));
    torch::optim::Adam *opts[6] = {m->meansOpt, m->scalesOpt, m->quatsOpt, m->featuresDcOpt, m->featuresRestOpt,
                                   m->opacitiesOpt};
    for (int i = 0; i < 6; ++i) inject_state(opts[i], adam_m[i], adam_v[i]);
    m->xys = torch::zeros_like(xys_grad).requires_grad_();
    m->xys.mutable_grad() = xys_grad.clone();
    m->radii = radii.clone();
    m->lastHeight = (int)height;
    m->lastWidth = (int)width;
    if (stats.size() == 3) {
        m->xysGradNorm = stats[0].clone();
        m->visCounts = stats[1].clone();
        m->max2DSize = stats[2].clone();
    }
    torch::manual_seed((uint64_t)seed);   // the stream Model::afterTrain's torch::randn (model.cpp:359) draws from
    m->afterTrain((int)step);
    std::vector<Tensor> out = {m->means.detach().clone(),      m->scales.detach().clone(),
                               m->quats.detach().clone(),      m->featuresDc.detach().clone(),
                               m->featuresRest.detach().clone(), m->opacities.detach().clone()};
    torch::optim::Adam *opts2[6] = {m->meansOpt, m->scalesOpt, m->quatsOpt, m->featuresDcOpt, m->featuresRestOpt,
                                    m->opacitiesOpt};
    std::vector<Tensor> ms, vs;
    for (int i = 0; i < 6; ++i) {
        auto mv = read_state(opts2[i]);
        ms.push_back(mv.first);
        vs.push_back(mv.second);
    }
    out.insert(out.end(), ms.begin(), ms.end());
    out.insert(out.end(), vs.begin(), vs.end());
    auto or_empty = [](const Tensor &t) { return t.defined() ? t.clone() : torch::empty({0}); };
    out.push_back(or_empty(m->xysGradNorm));
    out.push_back(or_empty(m->visCounts));
    out.push_back(or_empty(m->max2DSize));
    return out;
}

void save(std::vector<Tensor> params, std::string filename, int64_t step, bool keepCrs, double scale,
          Tensor translation) {
    auto m = make_model(params, 1, 100, 500, 30, 0.0002, 0.01, 4000, 0.05, 30000, keepCrs, scale, translation);
    m->save(filename, (int)step);
}

}  // namespace

TORCH_LIBRARY(opensplat_ref_model, m) {
    m.def("after_train", &after_train);
    m.def("save", &save);
}
