// stages.cpp -- TEST INFRASTRUCTURE (oracle/_ref/libsuma_ref_full.so): instantiates the generated adapters of every
// transpiled shader stage (oracle/_ref/gen/*.hpp under -DSGL_REFLECT) and registers them under the file name the
// reference passes to glow::GlShader::fromCache ("shader/gen_vertexmap.vert", ...). The list of headers is generated
// by glsl2cpp.py (all_stages.inc) from the shader list of ../Makefile.
#define SGL_REFLECT 1
#include "sgl.hpp"
#include "all_stages.inc"
