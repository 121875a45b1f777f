// stand-in for <glow/glexception.h>: see glow_all.hpp (TEST INFRASTRUCTURE, oracle/_ref)
#pragma once
#include "glow_all.hpp"
