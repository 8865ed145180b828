// frontend.cuh — launch interface of the range-image projection and LOAM feature-extraction kernels.
#pragma once
#include "common.cuh"
