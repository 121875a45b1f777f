// stand-in for <glow/GlTextureBuffer.h>: see glow_all.hpp (TEST INFRASTRUCTURE, oracle/_ref)
#pragma once
#include "glow_all.hpp"
