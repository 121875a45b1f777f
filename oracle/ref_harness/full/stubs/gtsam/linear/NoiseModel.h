// stand-in for <gtsam/linear/NoiseModel.h> (TEST INFRASTRUCTURE, oracle/_ref): only the names core/Posegraph.h mentions.
// The pose-graph optimiser (gtsam 4.0) is outside the hot path (out of scope) and not in this image.
#pragma once
#include <memory>
namespace gtsam {
namespace noiseModel { namespace mEstimator {
class Base { public: typedef std::shared_ptr<Base> shared_ptr; virtual ~Base() {} };
class DCS : public Base { public: static shared_ptr Create(double) { return shared_ptr(new DCS()); } };
class Huber : public Base { public: static shared_ptr Create(double) { return shared_ptr(new Huber()); } };
class Cauchy : public Base { public: static shared_ptr Create(double) { return shared_ptr(new Cauchy()); } };
class GemanMcClure : public Base { public: static shared_ptr Create(double) { return shared_ptr(new GemanMcClure()); } };
class Tukey : public Base { public: static shared_ptr Create(double) { return shared_ptr(new Tukey()); } };
class Welsh : public Base { public: static shared_ptr Create(double) { return shared_ptr(new Welsh()); } };
} }
class Values {};
class NonlinearFactorGraph {};
}  // namespace gtsam
