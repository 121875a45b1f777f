// stand-in for <glow/GlTransformFeedback.h>: see glow_all.hpp (TEST INFRASTRUCTURE, oracle/_ref)
#pragma once
#include "glow_all.hpp"
