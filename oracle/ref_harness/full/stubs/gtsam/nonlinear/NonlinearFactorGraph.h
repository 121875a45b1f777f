// stand-in for <gtsam/nonlinear/NonlinearFactorGraph.h> (TEST INFRASTRUCTURE, oracle/_ref): see ../linear/NoiseModel.h
#pragma once
#include "../linear/NoiseModel.h"
