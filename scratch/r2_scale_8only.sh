#!/bin/bash
NS="8" ./scratch/r2_scale_8gpu.sh
