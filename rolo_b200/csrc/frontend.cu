#include "frontend.cuh"
// front-end kernels land here (next commit)
#include "../../include/rolo_gicp.h"
extern "C" {
int rg_image_projection(rg_ctx*, const void*, size_t, const rg_projection_params*, float*, float*, int32_t*, float*, int32_t*, int32_t*, int32_t*) { return RG_ERR_UNSUPPORTED; }
int rg_feature_extraction(rg_ctx*, const float*, const int32_t*, const float*, const int32_t*, const int32_t*, int32_t, const rg_feature_params*, float*, int32_t*, int32_t*, float*, int32_t*, float*, int32_t*) { return RG_ERR_UNSUPPORTED; }
}
