// small_gicp/registration/reduction_omp.hpp — the include path the reference names (src/backMapping.cpp:5-10); everything lives in b200_adapter.hpp
#include <small_gicp/b200_adapter.hpp>
