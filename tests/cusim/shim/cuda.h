// TEST INFRASTRUCTURE: stands in for the CUDA header of the same name when the sources are built for the CPU executor
#pragma once
#include "../cusim.hpp"
