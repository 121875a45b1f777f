// stand-in for <glow/GlVertexArray.h>: see glow_all.hpp (TEST INFRASTRUCTURE, oracle/_ref)
#pragma once
#include "glow_all.hpp"
