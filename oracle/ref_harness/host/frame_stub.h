// frame_stub.h -- TEST INFRASTRUCTURE (oracle/_ref). core/Objective.h includes core/Frame.h, which holds OpenGL objects
// (glow, not vendored). The host harness compiles with -DINCLUDE_CORE_FRAME_H_ (Frame.h's own include guard) and this
// opaque stand-in: LieGaussNewton / Objective only pass std::shared_ptr<Frame> through.
#ifndef SUMA_REF_FRAME_STUB_H
#define SUMA_REF_FRAME_STUB_H
#include <memory>
#include <stdexcept>
class Frame {};
#endif
