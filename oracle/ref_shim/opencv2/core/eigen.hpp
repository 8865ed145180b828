// ref_shim: empty stand-in (the Scan Context sources include this header but use nothing from it)
#pragma once
#include <math.h>   // the real header chain exposes the float overloads of the C math functions in the global namespace
