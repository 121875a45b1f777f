// stand-in for <glow/GlTextureRectangle.h>: see glow_all.hpp (TEST INFRASTRUCTURE, oracle/_ref)
#pragma once
#include "glow_all.hpp"
