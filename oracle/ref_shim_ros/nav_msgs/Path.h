// forwards to the single stand-in header (test infrastructure only)
#include "rolo_ros_shim.h"
