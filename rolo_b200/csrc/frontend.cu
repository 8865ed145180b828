#include "frontend.cuh"
